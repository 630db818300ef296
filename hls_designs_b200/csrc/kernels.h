// Internal interface between the C-ABI layer (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include "../../include/hlsd.h"

namespace hlsd {

int option(int key);  // hlsd_set_option() value (0 = default)
void count_launch();  // bumps the process-wide launch counter reported by hlsd_launch_count()
int sm_count();
// cudaFuncAttributeMaxDynamicSharedMemorySize, issued once per (kernel, device) - again only if a later call needs
// more - and checked
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);
template <class K>
inline cudaError_t ensure_dynamic_smem(K* kernel, size_t bytes) {
    return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}
// Per-launch timing for hlsd_profile_begin/_end (include/hlsd.h): when profiling is on, a ProfScope brackets the
// launches issued during its lifetime with two CUDA events on their stream and files the pair under `cls`
// (HLSD_PROF_*); when it is off (the default) it costs one relaxed atomic load.
struct ProfScope {
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;
    int cls;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
};
// stream-ordered workspace memory from a library-owned pool (the device's default pool is left alone)
cudaError_t ws_alloc(void** p, size_t bytes, cudaStream_t st);
void ws_free(void* p, cudaStream_t st);

// Exact paths: every float operation of the reference design, in its per-element order, rounded
// separately.  `path` is one of HLSD_PATH_* (AUTO picks by size).  All enqueue on `st`.
// `fused` != 0 (HLSD_FLAG_FUSED): kernels that have a fused form contract a - b*c / the Givens pair into FMAs.
cudaError_t cholesky_exact(const float* A, float* L, int n, long batch, int variant, int path, int fused, cudaStream_t st);
cudaError_t lu_exact(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, int path, int fused,
                     cudaStream_t st);
cudaError_t qr_exact(const float* A, float* Q, float* R, int rows, int cols, long batch, int variant, int path, int fused,
                     cudaStream_t st);

// Blocked right-looking Cholesky for large N (n % 64 == 0, n >= 128): SIMT panel factorisation +
// trailing update on CUDA cores in the reference's per-element order (use_tensor == 0, exact) or
// on the tcgen05 tensor cores (use_tensor == 1).  variant 1 (cholesky_v2.2): exact form only.
// cudaErrorNotSupported when the size is not served.
cudaError_t cholesky_blocked(const float* A, float* L, int n, long batch, int use_tensor, int variant, cudaStream_t st);

// Bit-identical Cholesky (right-looking designs) of the m x m (m = 64, 128) diagonal block at (o, o) of matrices with row
// stride n, held in shared memory and blocked by 32 columns (blocked_exact.cu); n == m, o == 0: a batch of m x m matrices.
cudaError_t cholesky_tile_exact(const float* S, float* L, int m, int n, int o, long batch, int fused, cudaStream_t st);

// Bit-identical Cholesky (right-looking designs) of 128 x 128 matrices, register-tiled, one CTA per matrix.
cudaError_t cholesky128_exact(const float* A, float* L, long batch, int fused, cudaStream_t st);

// Blocked LU with partial pivoting on the tensor cores (n % 128 == 0, 256 <= n <= 1024): Crout order, rows never
// moved (lu_crout.cu).
cudaError_t lu_blocked(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, cudaStream_t st);

// The same factorisation in the reference's exact arithmetic and per-element order (bit-identical L, U, P).
cudaError_t lu_blocked_exact(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, cudaStream_t st);

// Self-check: shared-divisor division (common.cuh) vs IEEE division over `count` device-resident pairs;
// out[0] = bitwise mismatches, out[1] = pairs served by the fast path, out[2] = mismatches of the unguarded
// sequence (device memory, 3 x int64).
cudaError_t selftest_division(const float* num, const float* den, long count, long long* out2, cudaStream_t st);

}  // namespace hlsd
