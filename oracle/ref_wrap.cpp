// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI wrapper compiled TOGETHER WITH one reference design
// (its chol.cpp / lu.cpp / qr.cpp, taken where it lies under /root/reference or expanded from
// its template into oracle/_ref/gen) so that tests and bench.py's cpu_baseline leg can call the
// reference's own `top()` on arbitrary operands through ctypes.  Nothing here computes anything:
// it only casts flat row-major buffers to the reference's fixed-size 2-D array parameters and
// loops over a batch (optionally with OpenMP threads; the reference's top() keeps all its state
// in locals, so independent matrices may run concurrently).
#if defined(REF_CHOL)
#include "chol.h"
extern "C" int ref_rows() { return matrix_size; }
extern "C" int ref_cols() { return matrix_size; }
extern "C" int ref_chol(const float* A, float* L, long batch, int threads) {
    const long nn = (long)matrix_size * matrix_size;
    int rc = 0;
#pragma omp parallel for num_threads(threads) reduction(| : rc) schedule(static)
    for (long b = 0; b < batch; b++)
        rc |= top((matrix_t(*)[matrix_size])(A + b * nn), (matrix_t(*)[matrix_size])(L + b * nn));
    return rc;
}
#elif defined(REF_LU)
#include "lu.h"
extern "C" int ref_rows() { return matrix_size; }
extern "C" int ref_cols() { return matrix_size; }
extern "C" int ref_lu(const float* A, float* L, float* U, int* P, long batch, int threads) {
    const long nn = (long)matrix_size * matrix_size;
    int rc = 0;
#pragma omp parallel for num_threads(threads) reduction(| : rc) schedule(static)
    for (long b = 0; b < batch; b++) {
        uint_i p[matrix_size];
        rc |= top((matrix_t(*)[matrix_size])(A + b * nn), (matrix_t(*)[matrix_size])(L + b * nn),
                  (matrix_t(*)[matrix_size])(U + b * nn), p);
        for (int i = 0; i < matrix_size; i++) P[b * matrix_size + i] = (int)p[i];
    }
    return rc;
}
#elif defined(REF_QR)
#include "qr.h"
extern "C" int ref_rows() { return ROWS; }
extern "C" int ref_cols() { return COLS; }
extern "C" int ref_qr(const float* A, float* Q, float* R, long batch, int threads) {
    const long na = (long)ROWS * COLS, nq = (long)ROWS * ROWS;
    int rc = 0;
#pragma omp parallel for num_threads(threads) reduction(| : rc) schedule(static)
    for (long b = 0; b < batch; b++)
        rc |= top((MATRIX_T(*)[COLS])(A + b * na), (MATRIX_T(*)[ROWS])(Q + b * nq),
                  (MATRIX_T(*)[COLS])(R + b * na));
    return rc;
}
#else
#error "define REF_CHOL, REF_LU or REF_QR"
#endif
