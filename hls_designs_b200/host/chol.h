// Drop-in replacement for the reference's Cholesky/<design>/<size>/chol.h
// (Cholesky/cholesky_v1.3/model4x4/chol.h:1-18): same `matrix_t`, same
// `int top(matrix_t A[matrix_size][matrix_size], matrix_t L[matrix_size][matrix_size])`,
// computed on the GPU through the C ABI of libhlsd.so (include/hlsd.h) instead of by the
// synthesised systolic array.  Define matrix_size (the reference's headers hard-code it) and,
// optionally, HLSD_CHOL_DESIGN before including.
#ifndef HLSD_HOST_CHOL_H
#define HLSD_HOST_CHOL_H
#include "hlsd.h"
#ifndef matrix_size
#error "define matrix_size (e.g. -Dmatrix_size=16), as the reference's chol.h does"
#endif
#ifndef HLSD_CHOL_DESIGN
#define HLSD_CHOL_DESIGN HLSD_CHOL_V1_3
#endif
typedef float matrix_t;
static inline int top(matrix_t A[matrix_size][matrix_size], matrix_t L[matrix_size][matrix_size]) {
    return hlsd_cholesky_host(&A[0][0], &L[0][0], matrix_size, 1, HLSD_CHOL_DESIGN, HLSD_PATH_AUTO | HLSD_FLAG_EXACT);  // the csim's bits at every size
}
#endif
