// TEST CODE (not part of the product package).  Host-side testbench in the shape of the reference's
// *_tb.cpp, written against the drop-in headers of hls_designs_b200/host/: read op<N>.bin row by row with fread (chol_tb.cpp:18-24), call top()
// five times (chol_tb.cpp:25-28), rebuild the textbook factorisation in the testbench and compare
// (template/T_chol_tb.cpp:39-77, LU/lu_v1.1/model4x4/lu_tb.cpp:75-110 and :168-183,
// QR/qr_v2.1/template/T_qr_tb.cpp), return non-zero on failure.  The operand path is argv[1]
// instead of the author's absolute path.  Build: see hls_designs_b200/build.py (host_testbenches).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#if defined(TB_CHOL)
#include "chol.h"
static matrix_t A[matrix_size][matrix_size], A2[matrix_size][matrix_size], L1[matrix_size][matrix_size],
    L2[matrix_size][matrix_size];
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    for (int i = 0; i < matrix_size; i++)
        if (fread(A[i], sizeof(float), matrix_size, fp) != (size_t)matrix_size) return 2;
    fclose(fp);
    for (int i = 0; i < matrix_size; i++)
        for (int j = 0; j < matrix_size; j++) A2[i][j] = A[i][j], L2[i][j] = 0;
    int rc = 0;
    for (int i = 0; i < 5; i++) rc |= top(A, L1);
    if (rc) {
        fprintf(stderr, "top() failed: %s\n", hlsd_last_error());
        return 3;
    }
    for (int i = 0; i < matrix_size; i++) {  // the testbench's own right-looking loop
        L2[i][i] = sqrtf(A2[i][i]);
        for (int j = i + 1; j < matrix_size; j++) {
            L2[j][i] = A2[j][i] / L2[i][i];
            for (int k = i + 1; k <= j; k++) A2[j][k] = A2[j][k] - L2[j][i] * L2[k][i];
        }
    }
    float error = 0;
    for (int i = 0; i < matrix_size; i++)
        for (int j = 0; j < matrix_size; j++) error += fabsf(L1[i][j] - L2[i][j]);
    printf("error = %.6f\n", error);
    return error > 1e-5f ? 1 : 0;
}
#elif defined(TB_LU)
#include "lu.h"
static matrix_t A[matrix_size][matrix_size], A1[matrix_size][matrix_size], L[matrix_size][matrix_size],
    U[matrix_size][matrix_size], L1[matrix_size][matrix_size], U1[matrix_size][matrix_size];
static uint_i P[matrix_size], P1[matrix_size];
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    for (int i = 0; i < matrix_size; i++)
        if (fread(A[i], sizeof(float), matrix_size, fp) != (size_t)matrix_size) return 2;
    fclose(fp);
    for (int i = 0; i < matrix_size; i++)
        for (int j = 0; j < matrix_size; j++) A1[i][j] = A[i][j], U1[i][j] = 0, L1[i][j] = (i == j);
    int rc = 0;
    for (int i = 0; i < 5; i++) rc |= top(A, L, U, P);
    if (rc) {
        fprintf(stderr, "top() failed: %s\n", hlsd_last_error());
        return 3;
    }
    for (int i = 0; i < matrix_size - 1; i++) {  // the testbench's Doolittle loop with partial pivoting
        float maxpwr = fabsf(A1[i][i]);
        P1[i] = i;
        for (int j = i + 1; j < matrix_size; j++)
            if (maxpwr < fabsf(A1[j][i])) maxpwr = fabsf(A1[j][i]), P1[i] = j;
        if (P1[i] != (uint_i)i)
            for (int j = i; j < matrix_size; j++) {
                float t = A1[P1[i]][j];
                A1[P1[i]][j] = A1[i][j];
                A1[i][j] = t;
            }
        for (int j = i; j < matrix_size; j++) U1[i][j] = A1[i][j];
        for (int j = i + 1; j < matrix_size; j++) L1[j][i] = A1[j][i] / A1[i][i];
        for (int j = i + 1; j < matrix_size; j++)
            for (int k = i + 1; k < matrix_size; k++) A1[k][j] = A1[k][j] - L1[k][i] * U1[i][j];
    }
    P1[matrix_size - 1] = matrix_size - 1;
    U1[matrix_size - 1][matrix_size - 1] = A1[matrix_size - 1][matrix_size - 1];
    float error = 0;
    int perr = 0;
    for (int i = 0; i < matrix_size; i++) {
        perr += P[i] != P1[i];
        for (int j = 0; j < matrix_size; j++) error += fabsf(L[i][j] - L1[i][j]) + fabsf(U[i][j] - U1[i][j]);
    }
    printf("error = %.6f pivots differing = %d\n", error, perr);
    return (error > 5.0f || perr) ? 1 : 0;
}
#elif defined(TB_QR)
#include "qr.h"
static MATRIX_T A[ROWS][COLS], A1[ROWS][COLS], Q[ROWS][ROWS], R[ROWS][COLS], Q1[ROWS][ROWS];
static void mm(float c, float s, float& op1, float& op2) {
    float a = op2 * s + op1 * c, b = op2 * c - op1 * s;
    op1 = a, op2 = b;
}
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    for (int i = 0; i < ROWS; i++)
        if (fread(A[i], sizeof(float), COLS, fp) != (size_t)COLS) return 2;
    fclose(fp);
    for (int i = 0; i < ROWS; i++) {
        for (int j = 0; j < COLS; j++) A1[i][j] = A[i][j];
        for (int j = 0; j < ROWS; j++) Q1[i][j] = (i == j);
    }
    int rc = 0;
    for (int i = 0; i < 4; i++) rc |= top(A, Q, R);
    if (rc) {
        fprintf(stderr, "top() failed: %s\n", hlsd_last_error());
        return 3;
    }
    for (int i = 0; i < COLS; i++)  // the testbench's Givens loop (adjacent rows, bottom up)
        for (int j = ROWS - 1; j > i; j--) {
            if (A1[j][i] == 0) continue;
            float mag = sqrtf(A1[j][i] * A1[j][i] + A1[j - 1][i] * A1[j - 1][i]);
            float c = A1[j - 1][i] / mag, s = A1[j][i] / mag;
            A1[j - 1][i] = mag, A1[j][i] = 0;
            for (int k = i + 1; k < COLS; k++) mm(c, s, A1[j - 1][k], A1[j][k]);
            for (int k = 0; k < ROWS; k++) mm(c, s, Q1[k][j - 1], Q1[k][j]);
        }
    float error = 0;
    for (int i = 0; i < ROWS; i++) {
        for (int j = 0; j < ROWS; j++) error += fabsf(Q[i][j] - Q1[i][j]);
        for (int j = 0; j < COLS; j++) error += fabsf(R[i][j] - A1[i][j]);
    }
    printf("error = %.6f\n", error);
    return error > 5.0f ? 1 : 0;
}
#else
#error "define TB_CHOL, TB_LU or TB_QR"
#endif
