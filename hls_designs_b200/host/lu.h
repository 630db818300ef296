// Drop-in replacement for the reference's LU/<design>/<size>/lu.h (LU/lu_v1.1/model4x4/lu.h:1-24):
// `int top(A, L, U, P)` with the reference's output layout (unit-lower L, upper U, P[k] = row
// exchanged with row k).  The reference's `uint_i` is ap_uint<iterator_bit>; here it is a 32-bit
// unsigned integer with the same values.
#ifndef HLSD_HOST_LU_H
#define HLSD_HOST_LU_H
#include <stdint.h>
#include "hlsd.h"
#ifndef matrix_size
#error "define matrix_size (e.g. -Dmatrix_size=16), as the reference's lu.h does"
#endif
#ifndef HLSD_LU_DESIGN
#define HLSD_LU_DESIGN HLSD_LU_V1_1
#endif
typedef float matrix_t;
typedef uint32_t uint_i;
static inline int top(matrix_t A[matrix_size][matrix_size], matrix_t L[matrix_size][matrix_size],
                      matrix_t U[matrix_size][matrix_size], uint_i P[matrix_size]) {
    return hlsd_lu_host(&A[0][0], &L[0][0], &U[0][0], (int32_t*)P, matrix_size, 1, HLSD_LU_DESIGN, HLSD_PATH_AUTO | HLSD_FLAG_EXACT);  // the csim's bits at every size
}
#endif
