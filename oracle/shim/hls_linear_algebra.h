// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the pieces of Xilinx's hls_linear_algebra.h
// that the reference designs and testbenches use: x_sqrt, matrix_multiply, transpose tags.
#ifndef ORACLE_SHIM_HLS_LINEAR_ALGEBRA_H
#define ORACLE_SHIM_HLS_LINEAR_ALGEBRA_H
#include "hls_math.h"
namespace hls {
struct NoTranspose { static const bool T = false; };
struct Transpose { static const bool T = true; };
static inline float x_sqrt(float x) { return std::sqrt(x); }
static inline double x_sqrt(double x) { return std::sqrt(x); }
// C = op(A) * op(B), plain triple loop accumulating in OutT (testbench helper only).
template <class TA, class TB, int RA, int CA, int RB, int CB, int RC, int CC, typename InT, typename OutT>
void matrix_multiply(const InT A[RA][CA], const InT B[RB][CB], OutT C[RC][CC]) {
    const int K = TA::T ? RA : CA;
    for (int i = 0; i < RC; i++)
        for (int j = 0; j < CC; j++) {
            OutT acc = 0;
            for (int k = 0; k < K; k++) {
                InT a = TA::T ? A[k][i] : A[i][k];
                InT b = TB::T ? B[j][k] : B[k][j];
                acc += a * b;
            }
            C[i][j] = acc;
        }
}
}  // namespace hls
#endif
