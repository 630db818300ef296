/* hlsd.h — C ABI of libhlsd.so: batched Cholesky / LU / Givens-QR on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary for the top-level functions of Sibylau/HLS_designs.  Each reference design is
 * a fixed-size function over row-major float arrays:
 *
 *   int top(matrix_t A[N][N], matrix_t L[N][N]);                               Cholesky
 *       Cholesky/cholesky_v{1.3,2.2,3.2,4.0}/model4x4/chol.h:15-16
 *   int top(matrix_t A[N][N], matrix_t L[N][N], matrix_t U[N][N], uint_i P[N]); LU
 *       LU/lu_v{1.1,1.2,2.1}/model4x4/lu.h:19-22
 *   int top(MATRIX_T A[ROWS][COLS], MATRIX_T Q[ROWS][ROWS], MATRIX_T R[ROWS][COLS]); QR
 *       QR/qr_v{1.1,1.2,2.1}/model4x4/qr.h:22-24
 *
 * The entry points below are those functions with the compile-time size turned into arguments
 * and a leading batch dimension added: matrix b of a batch starts at base + b*rows*cols floats,
 * every matrix row-major exactly as `matrix_t A[N][N]` (and as the op<N>.bin / op<M>x<N>.bin
 * operand files once read with fread(A[i], sizeof(float), N, fp), e.g. chol_tb.cpp:18-24).
 * Output layouts are the reference collectors': L lower triangular with zeros above the diagonal
 * (unit diagonal for LU), U upper triangular with zeros below, P[k] = row exchanged with row k at
 * step k (P[N-1] = N-1), Q ROWSxROWS, R ROWSxCOLS.
 *
 * `*_host` functions take HOST pointers (like the reference's top()): they copy in, run, copy out
 * and return when the outputs are complete.  `*_dev` functions take DEVICE pointers and a
 * cudaStream_t (as void*) and only enqueue work.  All return 0 on success (like top()) or an
 * hlsd_status; hlsd_last_error() describes the last failure of the calling thread.
 * There is no CPU implementation behind this interface.
 */
#ifndef HLSD_H
#define HLSD_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HLSD_OK = 0,
    HLSD_ERR_INVALID = 1,     /* bad size / null pointer / unknown design or flag */
    HLSD_ERR_CUDA = 2,        /* a CUDA call failed; see hlsd_last_error() */
    HLSD_ERR_UNSUPPORTED = 3  /* valid request this build cannot serve (e.g. forced path does not fit) */
} hlsd_status;

/* Which reference design's arithmetic to reproduce (results are element-wise identical to that
 * design's C-simulation on the exact paths). */
typedef enum {
    HLSD_CHOL_V1_3 = 13, /* right-looking 1-D array            cholesky_v1.3/model4x4/chol.cpp:25-63   */
    HLSD_CHOL_V2_2 = 22, /* revised, roots taken at the end     cholesky_v2.2/model4x4/chol.cpp:28-207  */
    HLSD_CHOL_V3_2 = 32, /* 2-D array (== v1.3 arithmetic)      cholesky_v3.2/model4x4/chol.cpp:94-181  */
    HLSD_CHOL_V4_0 = 40  /* 1-D, other projection (== v1.3)     cholesky_v4.0/model4x4/chol.cpp:64-130  */
} hlsd_chol_design;

typedef enum {
    HLSD_LU_V1_1 = 11, /* Doolittle, partial pivoting, 1-D     lu_v1.1/model4x4/lu.cpp:23-92   */
    HLSD_LU_V1_2 = 12, /* same arithmetic, other projection    lu_v1.2/model4x4/lu.cpp:75-167  */
    HLSD_LU_V2_1 = 21, /* same arithmetic, 2-D array           lu_v2.1/model4x4/lu.cpp:50-155  */
    HLSD_LU_NOPIVOT = 1 /* this library's unpivoted variant: same elimination, P[k] = k always */
} hlsd_lu_design;

typedef enum {
    HLSD_QR_V1_1 = 11, /* 1-D array, rows paired (i-2, i)      qr_v1.1/model4x4/qr.cpp:45-308 (A != QR: reference defect) */
    HLSD_QR_V1_2 = 12, /* == v1.1 arithmetic                   qr_v1.2/model4x4/qr.cpp                                    */
    HLSD_QR_V2_1 = 21  /* 2-D array, adjacent rows (i-1, i)    qr_v2.1/template/T_qr.cpp:63-300                           */
} hlsd_qr_design;

/* flags: low byte selects the kernel family (0 = automatic by size). */
#define HLSD_PATH_AUTO 0
#define HLSD_PATH_TPM 1     /* the register-resident kernels (one thread or a few lanes per matrix):
                               N in {4, 8, 16, 32, 64}, Cholesky also 128; exact                      */
#define HLSD_PATH_GROUP 2   /* warp/CTA per matrix, shared-memory resident; exact                     */
#define HLSD_PATH_GLOBAL 3  /* one CTA per matrix in global memory: any N, exact, slow                */
#define HLSD_PATH_BLOCKED 4 /* blocked right-looking factorisation with an exact-order SIMT trailing update:
                               bit-identical to the csim like GLOBAL, 10-50x faster.  Cholesky: every design
                               (cholesky_v2.2 in its own arithmetic), n >= 256, n % 128 == 0; LU: every design incl. HLSD_LU_NOPIVOT,
                               256 <= n <= 1024, n % 128 == 0; 16-byte aligned buffers                        */
#define HLSD_PATH_TENSOR 5  /* blocked, tcgen05 tensor-core trailing update (3xTF32, fp32 accumulate):
                               Cholesky n >= 256, n % 128 == 0; LU 256 <= n <= 1024, n % 128 == 0;
                               16-byte aligned buffers.  NOT bit-identical to the csim: element-wise
                               agreement 1e-5 * max|L| (Cholesky) / 1e-3 * max|U| (LU, same pivots
                               unless two candidates tie to round-off), residuals < 2e-6 / 5e-5      */
#define HLSD_PATH_MASK 0xff
/* HLSD_PATH_AUTO picks the fastest kernel that serves the request: the register/shared-memory exact
 * kernels for n <= ~230; cholesky_v2.2 at n >= 256, n % 128 == 0 its blocked exact path; and for the right-looking
 * Cholesky designs / the LU designs at the
 * sizes and alignment listed under HLSD_PATH_TENSOR the tensor path (so AUTO at n = 256, 512, 1024 is
 * tolerance-exact, and an unaligned pointer or n = 255 stays bit-exact).  Callers that need the csim's
 * bits at every size set HLSD_FLAG_EXACT (AUTO then picks HLSD_PATH_BLOCKED where it applies, the register /
 * shared-memory / global kernels elsewhere); the drop-in top() headers (hls_designs_b200/host/) do. */
#define HLSD_FLAG_EXACT 0x100 /* never leave the reference design's arithmetic: AUTO does not select the
                                 tensor path.  With HLSD_PATH_TENSOR: HLSD_ERR_INVALID                 */
#define HLSD_FLAG_FUSED 0x200 /* opt-in speed mode of the register-resident kernels (n in 4..128): the
                                 reference's operations in the reference's order, but a - b*c and the
                                 Givens pair are contracted into fused multiply-adds (one rounding
                                 instead of two).  Not bit-identical; residuals as for the exact
                                 kernels; LU pivots may differ where two candidates tie to round-off.
                                 Ignored by kernels without a fused form.  With HLSD_FLAG_EXACT:
                                 HLSD_ERR_INVALID                                                      */
#define HLSD_FLAG_MASK 0x300

/* ---- host-pointer entry points (the reference's top(), batched) ---- */
int hlsd_cholesky_host(const float* A, float* L, int n, int64_t batch, int design, int flags);
int hlsd_lu_host(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags);
int hlsd_qr_host(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags);

/* ---- device-pointer entry points (enqueue on `stream`, a cudaStream_t; NULL = default stream) ---- */
int hlsd_cholesky_dev(const float* dA, float* dL, int n, int64_t batch, int design, int flags, void* stream);
int hlsd_lu_dev(const float* dA, float* dL, float* dU, int32_t* dP, int n, int64_t batch, int design, int flags,
                void* stream);
int hlsd_qr_dev(const float* dA, float* dQ, float* dR, int rows, int cols, int64_t batch, int design, int flags,
                void* stream);

/* ---- one host batch split over several GPUs of the box ----
 * BASELINE north_star: "the batch is partitioned across the 8 GPUs of one box by plain batch splitting, with no
 * NCCL on the hot path".  Same arguments and results as the *_host entry points; `devices` lists `ndevices` CUDA
 * device ordinals (NULL = all visible devices).  The batch is cut into contiguous, balanced spans, one per
 * device; each span runs the same chunked copy/compute pipeline on its own device from its own worker thread.
 * Returns the first failure (all spans are still waited for). */
int hlsd_cholesky_host_multi(const float* A, float* L, int n, int64_t batch, int design, int flags,
                             const int* devices, int ndevices);
int hlsd_lu_host_multi(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags,
                       const int* devices, int ndevices);
int hlsd_qr_host_multi(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags,
                       const int* devices, int ndevices);

/* ---- pinned host buffers for the *_host entry points (optional; pageable memory also works) ---- */
void* hlsd_host_alloc(size_t bytes);
void hlsd_host_free(void* p);
/* The *_host entry points keep their device staging buffers and streams in a process-wide pool (one set per
 * concurrent call and device, reused by later calls from any thread), and the blocked paths keep their
 * workspaces in a library-owned memory pool; this frees everything that is not in use. */
void hlsd_release_workspace(void);

/* ---- housekeeping ---- */
int hlsd_set_device(int device);           /* cudaSetDevice for the calling thread */
int hlsd_device_count(void);
int64_t hlsd_launch_count(void);           /* kernels launched by this library so far (process-wide) */
const char* hlsd_last_error(void);         /* thread-local, never NULL */
const char* hlsd_version(void);

/* ---- measurement switches (diagnostic).  Where two kernels serve the same request, the library runs the one
 * that measured faster on B200; these switches select the other one so that the choice can be re-measured
 * (scripts/ab_*.py).  Every setting computes the same results (bit-identical on the exact paths).  Process-wide;
 * returns HLSD_ERR_INVALID for an unknown option. */
#define HLSD_OPT_CHOL_SLAB 1    /* Cholesky register kernels: 0 = default (N = 32 dense slab, N = 64 packed lower-triangle
                                   slab), 1 = the other layout, 2 = N = 32 packed with 8 warps per SM                */
#define HLSD_OPT_RESERVED_2 2   /* (was the LU 32x32 kernel choice; the rows-in-registers kernel is no longer built)     */
#define HLSD_OPT_DIAG_KERNEL 3  /* tensor paths, 128x128 diagonal block: 0 = default, 1 = the other implementation   */
#define HLSD_OPT_LOCKSTEP 4     /* unrolled register kernels (Cholesky 32/64, LU 64/128, QR 16/32): 0 = default, 1 = the other
                                   of {free-running warps, one CTA-wide barrier per step}                             */
#define HLSD_OPT_QR_THREADS 5   /* QR 64x64 / 128x128: 0 = one set of threads does the R and the Q work (default),
                                   1 = separate R and Q threads (round 1)                                            */
#define HLSD_OPT_CHOL_TILE 6    /* exact Cholesky 64 / 128 (right-looking designs): 0 = shared-memory tile kernel blocked by
                                   32 columns (default), 1 = the register kernels of round 1, 2 = tile kernel with 64
                                   instead of 128 threads per 64x64 matrix                                             */
#define HLSD_OPT_COUNT 8
int hlsd_set_option(int option, int value);

/* ---- per-kernel-class timing of the blocked paths (diagnostic; used by bench.py for the roofline of the dominant
 * kernel).  Between hlsd_profile_begin() and hlsd_profile_end() every kernel launch of the library is bracketed by
 * CUDA events on its stream; hlsd_profile_end() waits for them and ADDS the elapsed milliseconds per class to
 * ms[0..n-1] (classes >= n are dropped) and the launch counts to launches[0..n-1] (may be NULL).  Process-wide; not
 * meant to stay on in production (two events per launch). */
#define HLSD_PROF_OTHER 0    /* zero fill, copies of the working matrix                                      */
#define HLSD_PROF_UPDATE 1   /* tcgen05 long-K updates: blk_ll2_kernel (Cholesky), lu_ll_kernel (LU)          */
#define HLSD_PROF_DIAG 2     /* SIMT diagonal block: Cholesky + inverse / inverse of the unit-lower LU block   */
#define HLSD_PROF_TRSM 3     /* tcgen05 panel solve: blk_mma_kernel (Cholesky), lu_trsm_kernel (LU)            */
#define HLSD_PROF_PANEL 4    /* SIMT pivoted panel factorisation of the blocked LU                            */
#define HLSD_PROF_PERMUTE 5  /* tensor LU: assembly of L from the multiplier store; exact blocked LU: exchanges */
#define HLSD_PROF_EXACT 6    /* the exact (register / shared-memory / global) kernels                          */
#define HLSD_PROF_CLASSES 7
int hlsd_profile_begin(void);
int hlsd_profile_end(double* ms, int64_t* launches, int n);

/* ---- self-checks (diagnostics; no counterpart in the reference) ----
 * The exact kernels divide several numerators by one pivot with a shared reciprocal (csrc/common.cuh).
 * This compares that division with IEEE division, bit for bit, over `count` pairs of DEVICE floats.
 * out3 (device, 3 x int64): [0] = mismatches (must be 0), [1] = pairs served by the fast sequence,
 * [2] = informational: mismatches the fast sequence alone would produce without the guard that sends divisors
 * with an all-ones mantissa to IEEE division. */
int hlsd_selftest_division(const float* num, const float* den, int64_t count, int64_t* out3, void* stream);
/* The staging pipeline of the *_host entry points with the kernel launch left out: `units` records of
 * in_bytes each go host -> device, `units` records of out_bytes each come back (their content is whatever the
 * staging buffers held).  Measures what the host <-> device links of the box allow for a given record shape;
 * bench.py reports the e2e leg as a fraction of it. */
int hlsd_selftest_copy_host(const void* in, void* out, size_t in_bytes, size_t out_bytes, int64_t units);

#ifdef __cplusplus
}
#endif
#endif /* HLSD_H */
