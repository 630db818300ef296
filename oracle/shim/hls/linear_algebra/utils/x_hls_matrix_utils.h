// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Xilinx's x_hls_matrix_utils.h (print_matrix).
#ifndef ORACLE_SHIM_X_HLS_MATRIX_UTILS_H
#define ORACLE_SHIM_X_HLS_MATRIX_UTILS_H
#include <cstdio>
#include "hls_linear_algebra.h"
namespace hls {
template <int R, int C, typename T, class Trans>
void print_matrix(T M[R][C], const char* prefix = "") {
    for (int i = 0; i < R; i++) {
        std::printf("%s|", prefix);
        for (int j = 0; j < C; j++) std::printf(" %12.6f", (double)(Trans::T ? M[j][i] : M[i][j]));
        std::printf(" |\n");
    }
}
}  // namespace hls
#endif
