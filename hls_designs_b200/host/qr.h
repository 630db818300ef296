// Drop-in replacement for the reference's QR/<design>/<size>/qr.h (QR/qr_v2.1/model4x4/qr.h:1-28):
// `int top(MATRIX_T A[ROWS][COLS], MATRIX_T Q[ROWS][ROWS], MATRIX_T R[ROWS][COLS])`.
// ROWS < COLS is rejected like the reference's top() (qr_v2.1/model4x4/qr.cpp:325-332), by status.
#ifndef HLSD_HOST_QR_H
#define HLSD_HOST_QR_H
#include "hlsd.h"
#if !defined(ROWS) || !defined(COLS)
#error "define ROWS and COLS (e.g. -DROWS=16 -DCOLS=16), as the reference's qr.h does"
#endif
#ifndef HLSD_QR_DESIGN
#define HLSD_QR_DESIGN HLSD_QR_V2_1
#endif
typedef float MATRIX_T;
static inline int top(MATRIX_T A[ROWS][COLS], MATRIX_T Q[ROWS][ROWS], MATRIX_T R[ROWS][COLS]) {
    return hlsd_qr_host(&A[0][0], &Q[0][0], &R[0][0], ROWS, COLS, 1, HLSD_QR_DESIGN, HLSD_PATH_AUTO | HLSD_FLAG_EXACT);  // the csim's bits at every size
}
#endif
