// Batched Cholesky (A = L L^T), the reference's `int top(matrix_t A[N][N], matrix_t L[N][N])`
// (Cholesky/cholesky_v*/model4x4/chol.h:15-16) over a batch of row-major matrices.
//
// Two arithmetic variants, each bit-identical to the reference designs it replaces:
//   variant 0  right-looking: cholesky_v1.3 (chol.cpp:25-63), v3.2 (chol.cpp:94-181), v4.0 (chol.cpp:64-130)
//   variant 1  "revised" root-free elimination + final scaling: cholesky_v2.2 (chol.cpp:28-207)
//
// Kernels:
//   chol_tpm_kernel<N>   thread-per-matrix, N in {4, 8}: each warp owns a shared-memory slab of
//                        32 padded matrix slots filled by 1-D bulk async copies (TMA engine), each
//                        thread factors its matrix entirely in registers and the slab is written
//                        back with bulk stores.  HBM-bound (2*N*N*4 bytes per matrix).
//   chol_group_kernel<G> G threads per matrix, matrix resident in shared memory: any N that fits.
//   chol_global_kernel   one CTA per matrix, working in place in global memory/L2: any N (slow, exact).
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

// ------------------------------------------------------------------------------------------------
// thread-per-matrix
// ------------------------------------------------------------------------------------------------
template <int N>
struct TpmCfg {
    static constexpr int kSlot = N * N + 4;  // floats; +16 B so that 8 consecutive slots hit 8 bank groups
    static constexpr int kSlabFloats = 32 * kSlot;
    static constexpr int kMatBytes = N * N * 4;
};

template <int N, int VARIANT, bool FUSED>
__device__ __forceinline__ void chol_regs(float (&a)[N * (N + 1) / 2]) {
#define T(i, j) ((i) * ((i) + 1) / 2 + (j))
    if constexpr (VARIANT == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            float d = xsqrt(a[T(k, k)]);
            a[T(k, k)] = d;
#pragma unroll
            for (int i = k + 1; i < N; i++) a[T(i, k)] = xdiv(a[T(i, k)], d);
#pragma unroll
            for (int i = k + 1; i < N; i++)
#pragma unroll
                for (int j = i; j < N; j++) a[T(j, i)] = msub<FUSED>(a[T(j, i)], a[T(j, k)], a[T(i, k)]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) {
            float d = a[T(k, k)];
            float m[N];
#pragma unroll
            for (int i = k + 1; i < N; i++) m[i] = xdiv(a[T(i, k)], d);
#pragma unroll
            for (int i = k + 1; i < N; i++)
#pragma unroll
                for (int j = i; j < N; j++) a[T(j, i)] = msub<FUSED>(a[T(j, i)], a[T(j, k)], m[i]);
#pragma unroll
            for (int i = k + 1; i < N; i++) a[T(i, k)] = m[i];
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
            float r = xsqrt(a[T(k, k)]);
            a[T(k, k)] = r;
#pragma unroll
            for (int j = k + 1; j < N; j++) a[T(j, k)] = xmul(a[T(j, k)], r);
        }
    }
#undef T
}

template <int N, int VARIANT, int WARPS, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, 1)
chol_tpm_kernel(const float* __restrict__ A, float* __restrict__ L, long batch) {
    using Cfg = TpmCfg<N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * Cfg::kSlabFloats * 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* slab = slabs + (size_t)warp * Cfg::kSlabFloats;
    float* slot = slab + lane * Cfg::kSlot;
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const long num_slabs = (batch + 31) / 32;
    uint32_t parity = 0;
    for (long s = (long)blockIdx.x * WARPS + warp; s < num_slabs; s += (long)gridDim.x * WARPS) {
        const long m0 = s * 32;
        const int valid = (int)min(32L, batch - m0);
        constexpr bool kCoalesced = N <= 4;  // 64-byte matrices: cooperative coalesced copies beat one bulk copy per matrix
        if constexpr (kCoalesced) {
            __syncwarp();  // the previous slab has been written out of the slots
            slab_g2s<N * N / 4>(reinterpret_cast<float4*>(slab), reinterpret_cast<const float4*>(A + m0 * (long)(N * N)), valid, lane);
            __syncwarp();
        } else {
            // the bulk store that last read this slot (issued by this very thread) must have drained
            bulk_wait_read<0>();
            if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * Cfg::kMatBytes);
            __syncwarp();
            if (lane < valid) bulk_g2s(slot, A + (m0 + lane) * (long)(N * N), Cfg::kMatBytes, bar);
            mbar_wait(bar, parity);
            parity ^= 1;
        }
        if (lane < valid) {
            float a[N * (N + 1) / 2];
            const float4* s4 = reinterpret_cast<const float4*>(slot);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int q = 0; q <= i / 4; q++) {
                    float4 v = s4[i * (N / 4) + q];
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (4 * q + t <= i) a[i * (i + 1) / 2 + 4 * q + t] = e[t];
                }
            chol_regs<N, VARIANT, FUSED>(a);
            float4* d4 = reinterpret_cast<float4*>(slot);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int q = 0; q < N / 4; q++) {
                    float e[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) e[t] = (4 * q + t <= i) ? a[i * (i + 1) / 2 + 4 * q + t] : 0.0f;
                    d4[i * (N / 4) + q] = make_float4(e[0], e[1], e[2], e[3]);
                }
            if constexpr (!kCoalesced) {
                fence_async_smem();
                bulk_s2g(L + (m0 + lane) * (long)(N * N), slot, Cfg::kMatBytes);
                bulk_commit();
            }
        }
        if constexpr (kCoalesced) {
            __syncwarp();
            slab_s2g<N * N / 4>(reinterpret_cast<float4*>(L + m0 * (long)(N * N)), reinterpret_cast<const float4*>(slab), valid, lane);
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// R lanes per matrix (R = 2 or 4), rows dealt round-robin to the lanes of a group: row i lives in
// lane (i % R) as local row i / R.  Same bulk-copy slab staging as above, but a warp holds 32/R
// matrices, so slabs and register files are R times smaller and R times as many warps fit on an
// SM: the kernel is occupancy-bound (measured: 4/5/6 warps per SM -> 56/70/75 % of the HBM
// roofline with one thread per matrix).  Per step k the owner lane roots the pivot and broadcasts
// it by shuffle (the systolic "pivot broadcast"), every lane scales its own entries of column k,
// and the column is all-gathered by one shuffle per row (the "column update" stream).
// Entries a[t][j] with j > row index are scratch (computed, never stored).
// ------------------------------------------------------------------------------------------------
// PACKED: the slab holds only the lower triangle, row i rounded up to whole 16-byte pieces (i/4 + 1 of them), so a
// 32 x 32 matrix takes 2.3 KB instead of 4 KB and twice as many warps fit on an SM (the kernel is latency bound:
// 6 warps per SM gave 0.47 of the HBM roofline).  Every row arrives by its own bulk copy (only the pieces that
// hold a_ij, j <= i: 44 % less operand traffic too) and leaves by one bulk store for its data pieces and one, from
// a shared block of zeros, for the pieces right of the diagonal.
template <int N>
struct PackedTri {
    static constexpr int G = N / 4;
    static constexpr int kPieces = 2 * G * (G + 1);           // 16-byte pieces per matrix
    static constexpr int kSlot = (kPieces + 1) * 4;           // floats; +1 piece: consecutive slots start in different banks
    __host__ __device__ static constexpr int row_off(int i) {  // first piece of row i
        return (i / 4 + 1) * (2 * (i / 4) + i % 4);
    }
};

// LOCKSTEP: one CTA-wide barrier per elimination step.  The step loop is unrolled at compile time (register arrays
// can only be indexed by constants), so a 64 x 64 factorisation is ~190 KB of straight-line code that every warp
// walks through once per slab; with the warps of an SM at different places in it, instruction fetch becomes the
// top stall (ncu: `no_instruction` 3.6 warps per issue at N = 64).  In lockstep all warps fetch the same lines at
// the same time.  Needs the same number of slab iterations in every warp of a CTA (idle ones run on `valid = 0`).
template <int N, int R, int VARIANT, int WARPS, bool FUSED, bool PACKED = false, bool LOCKSTEP = false>
__global__ void __launch_bounds__(WARPS * 32, 1)
chol_rows_kernel(const float* __restrict__ A, float* __restrict__ L, long batch) {
    using Cfg = TpmCfg<N>;
    using Tri = PackedTri<N>;
    constexpr int MPW = 32 / R;  // matrices per warp
    constexpr int T = N / R;     // local rows per lane
    constexpr int kSlotFloats = PACKED ? Tri::kSlot : Cfg::kSlot;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * MPW * kSlotFloats * 4);
    __shared__ __align__(16) float zeros[PACKED ? N : 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane % R, m = lane / R, gbase = lane - r;
    float* slot = slabs + ((size_t)warp * MPW + m) * kSlotFloats;
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if constexpr (PACKED) {
        if (threadIdx.x < N) zeros[threadIdx.x] = 0.0f;
        fence_async_smem();
        __syncthreads();
    }
    __syncwarp();
    const long num_slabs = (batch + MPW - 1) / MPW;
    constexpr uint32_t kPart = Cfg::kMatBytes / R;  // dense layout: each lane of a group moves 1/R of the matrix
    uint32_t parity = 0;
    for (long s0 = (long)blockIdx.x * WARPS; s0 < num_slabs; s0 += (long)gridDim.x * WARPS) {
        const long s = s0 + warp;
        if (!LOCKSTEP && s >= num_slabs) break;
        const long m0 = s * MPW;
        const int valid = s < num_slabs ? (int)min((long)MPW, batch - m0) : 0;
        bulk_wait_read<0>();
        __syncwarp();  // all lanes' previous stores have drained before anyone refills the slab
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * (PACKED ? Tri::kPieces * 16 : Cfg::kMatBytes));
        __syncwarp();
        // a[t][j]: row i = t*R + r, columns 0 .. t*R + R - 1 (static bound; j > i is scratch)
        float a[T][N];
        if constexpr (PACKED) {
            if (m < valid) {
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int i = t * R + r;
                    bulk_g2s(slot + Tri::row_off(i) * 4, A + (m0 + m) * (long)(N * N) + i * N, (uint32_t)(i / 4 + 1) * 16, bar);
                }
            }
            mbar_wait(bar, parity);
            parity ^= 1;
#pragma unroll
            for (int t = 0; t < T; t++) {
                const int i = t * R + r;
                const float4* s4 = reinterpret_cast<const float4*>(slot) + Tri::row_off(i);
#pragma unroll
                for (int q = 0; q < (t * R + R + 3) / 4; q++) {
                    // pieces right of the diagonal piece are not in the slab: those registers stay scratch
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (R == 4 || q <= i / 4) v = s4[q];
                    a[t][4 * q] = v.x, a[t][4 * q + 1] = v.y, a[t][4 * q + 2] = v.z, a[t][4 * q + 3] = v.w;
                }
            }
        } else {
            // each lane moves 1/R of its matrix: many small copies in flight measured faster than one per matrix
            if (m < valid) bulk_g2s(slot + r * (kPart / 4), A + (m0 + m) * (long)(N * N) + r * (kPart / 4), kPart, bar);
            mbar_wait(bar, parity);
            parity ^= 1;
            const float4* s4 = reinterpret_cast<const float4*>(slot);
#pragma unroll
            for (int t = 0; t < T; t++) {
#pragma unroll
                for (int q = 0; q < (t * R + R + 3) / 4; q++) {
                    float4 v = s4[(t * R + r) * (N / 4) + q];
                    a[t][4 * q] = v.x, a[t][4 * q + 1] = v.y, a[t][4 * q + 2] = v.z, a[t][4 * q + 3] = v.w;
                }
            }
        }
        float dsave[VARIANT == 1 ? N : 1];
        static_for<0, N>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr int tk = k / R, rk = k % R;
            if constexpr (LOCKSTEP) __syncthreads();
            // pivot: owner lane rk, local row tk
            // lanes that do not own row k hold scratch there (possibly zero or negative): they root 1.0 instead,
            // so that no lane leaves the fast path of the square root
            float piv = (r == rk) ? a[tk][k] : 1.0f;
            if constexpr (VARIANT == 0) piv = xsqrt(piv);
            const float d = __shfl_sync(0xffffffffu, piv, gbase + rk);
            if constexpr (VARIANT == 0) {
                if (r == rk) a[tk][k] = d;
            } else {
                dsave[k] = d;
            }
            // own entries of column k below the diagonal: rows i = t*R + r > k
            float own[T];
            {
                const SharedDivisor dv = make_divisor(d);
                // local row tk is scratch in lanes r <= rk (rows <= k): divide the pivot by itself there
                float num[T];
                bool ok = dv.ok;
#pragma unroll
                for (int t = tk; t < T; t++) {
                    num[t] = (t > tk || r > rk) ? a[t][k] : d;
                    ok &= div_in_range(num[t]);
                }
                if (ok) {
#pragma unroll
                    for (int t = tk; t < T; t++) own[t] = fast_quotient(num[t], dv);
                } else {
#pragma unroll
                    for (int t = tk; t < T; t++) own[t] = xdiv(num[t], d);
                }
            }
            if constexpr (VARIANT == 0) {
#pragma unroll
                for (int t = tk; t < T; t++)
                    if (t > tk || r > rk) a[t][k] = own[t];
            }
            // all-gather column k (rows k+1 .. N-1) and update
            if constexpr (N >= 32) {
                // two columns per instruction (FMUL2 / FFMA2, common.cuh): columns (i, i+1), i even, have the same row
                // range t >= i/R (R is even); an odd first column goes alone
                constexpr int i0 = (k + 1) | 1;  // first odd column >= k+1 ... pairs start at the even column after it
                if constexpr ((k + 1) % 2 == 1 && k + 1 < N) {
                    constexpr int i = k + 1;
                    const float lik = __shfl_sync(0xffffffffu, own[i / R], gbase + (i % R));
#pragma unroll
                    for (int t = i / R; t < T; t++) a[t][i] = msub<FUSED>(a[t][i], a[t][k], lik);
                }
                constexpr int ie = ((k + 1) % 2 == 1) ? k + 2 : k + 1;  // even
                (void)i0;
#pragma unroll
                for (int i = ie; i + 1 < N; i += 2) {
                    const float l0 = __shfl_sync(0xffffffffu, own[i / R], gbase + (i % R));
                    const float l1 = __shfl_sync(0xffffffffu, own[(i + 1) / R], gbase + ((i + 1) % R));
#pragma unroll
                    for (int t = i / R; t < T; t++) msub_pair<FUSED>(a[t][i], a[t][i + 1], a[t][k], l0, l1);
                }
            } else {
#pragma unroll
            for (int i = k + 1; i < N; i++) {
                const float lik = __shfl_sync(0xffffffffu, own[i / R], gbase + (i % R));
                // rows j = t*R + r >= i  <=>  t >= i/R (the t == i/R case is scratch when r < i % R)
#pragma unroll
                for (int t = i / R; t < T; t++) a[t][i] = msub<FUSED>(a[t][i], a[t][k], lik);
            }
            }
            if constexpr (VARIANT == 1) {
#pragma unroll
                for (int t = tk; t < T; t++)
                    if (t > tk || r > rk) a[t][k] = own[t];
            }
        });
        if constexpr (VARIANT == 1) {
#pragma unroll
            for (int k = 0; k < N; k++) {
                const float rt = xsqrt(dsave[k]);
                const int tk = k / R, rk = k % R;
#pragma unroll
                for (int t = tk; t < T; t++) {
                    if (t > tk || r > rk) a[t][k] = xmul(a[t][k], rt);
                    else if (t == tk && r == rk) a[t][k] = rt;
                }
            }
        }
        if constexpr (PACKED) {
            if (m < valid) {
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int i = t * R + r;
                    float4* d4 = reinterpret_cast<float4*>(slot) + Tri::row_off(i);
#pragma unroll
                    for (int q = 0; q < (t * R + R + 3) / 4; q++) {
                        float e[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int j = 4 * q + u;
                            e[u] = (j < t * R + R && j <= i) ? a[t][j < t * R + R ? j : 0] : 0.0f;
                        }
                        if (R == 4 || q <= i / 4) d4[q] = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                fence_async_smem();
                // every lane stores the rows it wrote itself: the data pieces, then zeros for the rest of the row
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int i = t * R + r;
                    float* dst = L + (m0 + m) * (long)(N * N) + i * N;
                    const int pieces = i / 4 + 1;
                    bulk_s2g(dst, slot + Tri::row_off(i) * 4, (uint32_t)pieces * 16);
                    if (pieces < N / 4) bulk_s2g(dst + pieces * 4, zeros, (uint32_t)(N / 4 - pieces) * 16);
                }
                bulk_commit();
            }
        } else {
            if (m < valid) {
                float4* d4 = reinterpret_cast<float4*>(slot);
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int i = t * R + r;
#pragma unroll
                    for (int q = 0; q < N / 4; q++) {
                        float e[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int j = 4 * q + u;
                            e[u] = (j < t * R + R && j <= i) ? a[t][j < t * R + R ? j : 0] : 0.0f;
                        }
                        d4[i * (N / 4) + q] = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                fence_async_smem();
            }
            __syncwarp();  // the whole matrix (all R lanes' rows) is in the slot
            if (m < valid) {
                bulk_s2g(L + (m0 + m) * (long)(N * N) + r * (kPart / 4), slot + r * (kPart / 4), kPart);
                bulk_commit();
            }
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// G threads per matrix, matrix resident in shared memory
// ------------------------------------------------------------------------------------------------
template <int G, int VARIANT>
__global__ void __launch_bounds__(256)
chol_group_kernel(const float* __restrict__ A, float* __restrict__ L, int n, long batch) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int kGroups = 256 / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ld = n | 1;  // odd leading dimension: column walks do not collide on one bank
    float* w = smem_f + (size_t)g * ((size_t)n * ld + n);
    float* mcol = w + (size_t)n * ld;  // variant 1: multipliers of the current stage
    const long nn = (long)n * n;
    for (long b = (long)blockIdx.x * kGroups + g; b < batch; b += (long)gridDim.x * kGroups) {
        const float* a = A + b * nn;
        for (int e = t; e < n * n; e += G) {
            int i = e / n, j = e - i * n;
            w[i * ld + j] = a[e];
        }
        group_sync<G>(g);
        for (int k = 0; k < n; k++) {
            if constexpr (VARIANT == 0) {
                float d = xsqrt(w[k * ld + k]);
                group_sync<G>(g);  // everyone has read the un-rooted pivot
                if (t == 0) w[k * ld + k] = d;
                for (int i = k + 1 + t; i < n; i += G) w[i * ld + k] = xdiv(w[i * ld + k], d);
                group_sync<G>(g);
                // a[j][i] -= l[j][k] * l[i][k], k < i <= j: a warp-wide strip of row j per inner loop
                {
                    constexpr int GY = G / 32;
                    const int tx = t & 31, ty = t >> 5;
                    for (int j = k + 1 + ty; j < n; j += GY) {
                        float* wrow = w + j * ld;
                        const float ljk = wrow[k];
                        for (int i = k + 1 + tx; i <= j; i += 32) wrow[i] = xmsub(wrow[i], ljk, w[i * ld + k]);
                    }
                }
                group_sync<G>(g);
            } else {
                float d = w[k * ld + k];
                for (int i = k + 1 + t; i < n; i += G) mcol[i] = xdiv(w[i * ld + k], d);
                group_sync<G>(g);
                {
                    constexpr int GY = G / 32;
                    const int tx = t & 31, ty = t >> 5;
                    for (int j = k + 1 + ty; j < n; j += GY) {
                        float* wrow = w + j * ld;
                        const float ajk = wrow[k];
                        for (int i = k + 1 + tx; i <= j; i += 32) wrow[i] = xmsub(wrow[i], ajk, mcol[i]);
                    }
                }
                group_sync<G>(g);
                for (int i = k + 1 + t; i < n; i += G) w[i * ld + k] = mcol[i];
                group_sync<G>(g);
            }
        }
        if constexpr (VARIANT == 1) {
            for (int k = t; k < n; k += G) mcol[k] = xsqrt(w[k * ld + k]);
            group_sync<G>(g);
        }
        float* l = L + b * nn;
        for (int e = t; e < n * n; e += G) {
            int i = e / n, j = e - i * n;
            float v = 0.0f;
            if (j <= i) {
                v = w[i * ld + j];
                if constexpr (VARIANT == 1) v = (i == j) ? mcol[j] : xmul(v, mcol[j]);
            }
            l[e] = v;
        }
        group_sync<G>(g);
    }
}

// ------------------------------------------------------------------------------------------------
// one CTA per matrix, in place in global memory (any N; exact; slow)
// ------------------------------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(1024)
chol_global_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ scratch, int n, long batch) {
    const int t = threadIdx.x, G = blockDim.x;
    const long nn = (long)n * n;
    for (long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float* a = A + b * nn;
        float* w = L + b * nn;
        float* mcol = scratch + (long)blockIdx.x * n;
        for (long e = t; e < nn; e += G) w[e] = a[e];
        __syncthreads();
        for (int k = 0; k < n; k++) {
            const int m = n - k - 1;
            if constexpr (VARIANT == 0) {
                float d = xsqrt(w[(long)k * n + k]);
                __syncthreads();
                if (t == 0) w[(long)k * n + k] = d;
                for (int i = k + 1 + t; i < n; i += G) w[(long)i * n + k] = xdiv(w[(long)i * n + k], d);
                __syncthreads();
                for (long e = t; e < (long)m * m; e += G) {
                    int jj = (int)(e / m), ii = (int)(e - (long)jj * m);
                    if (ii <= jj) {
                        long j = k + 1 + jj, i = k + 1 + ii;
                        w[j * n + i] = xmsub(w[j * n + i], w[j * n + k], w[i * n + k]);
                    }
                }
                __syncthreads();
            } else {
                float d = w[(long)k * n + k];
                for (int i = k + 1 + t; i < n; i += G) mcol[i] = xdiv(w[(long)i * n + k], d);
                __syncthreads();
                for (long e = t; e < (long)m * m; e += G) {
                    int jj = (int)(e / m), ii = (int)(e - (long)jj * m);
                    if (ii <= jj) {
                        long j = k + 1 + jj, i = k + 1 + ii;
                        w[j * n + i] = xmsub(w[j * n + i], w[j * n + k], mcol[i]);
                    }
                }
                __syncthreads();
                for (int i = k + 1 + t; i < n; i += G) w[(long)i * n + k] = mcol[i];
                __syncthreads();
            }
        }
        if constexpr (VARIANT == 1) {
            for (int k = t; k < n; k += G) mcol[k] = xsqrt(w[(long)k * n + k]);
            __syncthreads();
        }
        for (long e = t; e < nn; e += G) {
            int i = (int)(e / n), j = (int)(e - (long)i * n);
            float v = 0.0f;
            if (j <= i) {
                v = w[e];
                if constexpr (VARIANT == 1) v = (i == j) ? mcol[j] : xmul(v, mcol[j]);
            }
            w[e] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
template <int N, int VARIANT, int WARPS, bool FUSED>
static cudaError_t launch_tpm_w(const float* A, float* L, long batch, cudaStream_t st) {
    using Cfg = TpmCfg<N>;
    const size_t smem = (size_t)WARPS * Cfg::kSlabFloats * 4 + WARPS * sizeof(uint64_t);
    auto kern = chol_tpm_kernel<N, VARIANT, WARPS, FUSED>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + 31) / 32;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, L, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, int R, int VARIANT, int WARPS, bool FUSED, bool PACKED = false, bool LOCKSTEP = false>
static cudaError_t launch_rows(const float* A, float* L, long batch, cudaStream_t st) {
    using Cfg = TpmCfg<N>;
    constexpr int MPW = 32 / R;
    constexpr int kSlotFloats = PACKED ? PackedTri<N>::kSlot : Cfg::kSlot;
    constexpr size_t smem = (size_t)WARPS * MPW * kSlotFloats * 4 + WARPS * sizeof(uint64_t);
    static_assert(smem <= 227 * 1024, "slabs do not fit");
    auto kern = chol_rows_kernel<N, R, VARIANT, WARPS, FUSED, PACKED, LOCKSTEP>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + MPW - 1) / MPW;
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count());
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, L, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, int VARIANT, bool FUSED>
static cudaError_t launch_tpm(const float* A, float* L, long batch, cudaStream_t st) {
    if constexpr (N == 16) {
        // lanes per matrix, measured at batch 1M (GB/s of algorithmic traffic): 1 (6 warps/SM) 4739, 2 (12 warps) 5477,
        // 4 (24 warps) 6017 -> 6305 after the division work of DESIGN.md 4.0
        return launch_rows<N, 4, VARIANT, 24, FUSED>(A, L, batch, st);
    } else if constexpr (N == 32) {
        // measured (fraction of HBM peak): 4 lanes per matrix 0.47, 8 lanes 0.39, 16 lanes 0.31: fewer lanes per matrix
        // mean fewer shuffles per useful operation, even at half the warps per SM
        // Packed lower-triangle slabs (12 warps per SM instead of 6) measured SLOWER here: 0.37 of the HBM roofline with
        // 12 warps (168 registers, ~170 words spilled), 0.33 with 8 warps (no spills), against 0.465 for the dense
        // slab: a row-wise packed slab costs 23 bulk copies of 16..128 bytes per lane and matrix instead of 2 of 1 KB.
        // (At N = 64 the same change is 6 copies per lane and wins: 0.181 vs 0.155.)
        if (option(HLSD_OPT_CHOL_SLAB) == 1) return launch_rows<N, 4, VARIANT, 12, FUSED, true>(A, L, batch, st);
        if (option(HLSD_OPT_CHOL_SLAB) == 2) return launch_rows<N, 4, VARIANT, 8, FUSED, true>(A, L, batch, st);
        // one CTA-wide barrier per step (LOCKSTEP) measured +5 % here (0.468 -> 0.490), -1 % at N = 64
        if (option(HLSD_OPT_LOCKSTEP) == 1) return launch_rows<N, 4, VARIANT, 6, FUSED>(A, L, batch, st);
        return launch_rows<N, 4, VARIANT, 6, FUSED, false, true>(A, L, batch, st);
    } else if constexpr (N == 64) {
        // (16 lanes per matrix measured slower than 32: 0.152 vs 0.168)
        if (option(HLSD_OPT_CHOL_SLAB) == 1) return launch_rows<N, 32, VARIANT, 12, FUSED>(A, L, batch, st);
        // packed slabs: 16 warps per SM instead of 12 (20 would cap the kernel at 96 registers: 2 KB of spills)
        if (option(HLSD_OPT_LOCKSTEP) == 1) return launch_rows<N, 32, VARIANT, 16, FUSED, true, true>(A, L, batch, st);
        return launch_rows<N, 32, VARIANT, 16, FUSED, true>(A, L, batch, st);
    } else {
        return launch_tpm_w<N, VARIANT, 8, FUSED>(A, L, batch, st);
    }
}

template <int G, int VARIANT>
static cudaError_t launch_group(const float* A, float* L, int n, long batch, cudaStream_t st) {
    constexpr int kGroups = 256 / G;
    const size_t per = ((size_t)n * (n | 1) + n) * 4;
    const size_t smem = per * kGroups;
    auto kern = chol_group_kernel<G, VARIANT>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((batch + kGroups - 1) / kGroups, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, 256, smem, st>>>(A, L, n, batch);
    count_launch();
    return cudaGetLastError();
}

template <int VARIANT, bool FUSED>
static cudaError_t chol_dispatch(const float* A, float* L, int n, long batch, int path, cudaStream_t st) {
    const bool aligned = (((uintptr_t)A | (uintptr_t)L) & 15) == 0;
    // 64 x 64 and 128 x 128, right-looking designs: the shared-memory tile kernel blocked by 32 columns (blocked_exact.cu);
    // HLSD_OPT_CHOL_TILE = 1: the register kernels (N = 128: the register-tiled CTA-per-matrix kernel of blocked.cu)
    if (VARIANT == 0 && aligned && (n == 64 || n == 128) && (path == HLSD_PATH_AUTO || path == HLSD_PATH_TPM) &&
        option(HLSD_OPT_CHOL_TILE) != 1)
        return cholesky_tile_exact(A, L, n, n, 0, batch, FUSED, st);
    if (VARIANT == 0 && n == 128 && (path == HLSD_PATH_AUTO || path == HLSD_PATH_TPM)) return cholesky128_exact(A, L, batch, FUSED, st);
    if (path == HLSD_PATH_AUTO || path == HLSD_PATH_TPM) {
        if (aligned && n == 4) return launch_tpm<4, VARIANT, FUSED>(A, L, batch, st);
        if (aligned && n == 8) return launch_tpm<8, VARIANT, FUSED>(A, L, batch, st);
        if (aligned && n == 16) return launch_tpm<16, VARIANT, FUSED>(A, L, batch, st);
        if (aligned && n == 32) return launch_tpm<32, VARIANT, FUSED>(A, L, batch, st);
        if (aligned && n == 64) return launch_tpm<64, VARIANT, FUSED>(A, L, batch, st);
        if (path == HLSD_PATH_TPM) return cudaErrorInvalidValue;
    }
    // (the kernels below have no fused form: HLSD_FLAG_FUSED is a permission, not a request)
    const size_t per = ((size_t)n * (n | 1) + n) * 4;
    const size_t cap = 227 * 1024 - 1024;
    if ((path == HLSD_PATH_AUTO && per <= cap) || path == HLSD_PATH_GROUP) {
        if (per > cap) return cudaErrorInvalidValue;
        if (n <= 48 && per * 8 <= cap) return launch_group<32, VARIANT>(A, L, n, batch, st);
        if (n <= 96 && per * 2 <= cap) return launch_group<128, VARIANT>(A, L, n, batch, st);
        return launch_group<256, VARIANT>(A, L, n, batch, st);
    }
    // global-memory fallback
    float* scratch = nullptr;
    long grid = std::min<long>(batch, (long)sm_count() * 2);
    cudaError_t e = ws_alloc(reinterpret_cast<void**>(&scratch), sizeof(float) * (size_t)grid * n, st);
    if (e != cudaSuccess) return e;
    chol_global_kernel<VARIANT><<<(unsigned)grid, 1024, 0, st>>>(A, L, scratch, n, batch);
    count_launch();
    e = cudaGetLastError();
    ws_free(scratch, st);
    return e;
}

cudaError_t cholesky_exact(const float* A, float* L, int n, long batch, int variant, int path, int fused, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    ProfScope ps(HLSD_PROF_EXACT, st);
    if (fused)
        return variant == 1 ? chol_dispatch<1, true>(A, L, n, batch, path, st) : chol_dispatch<0, true>(A, L, n, batch, path, st);
    return variant == 1 ? chol_dispatch<1, false>(A, L, n, batch, path, st) : chol_dispatch<0, false>(A, L, n, batch, path, st);
}

}  // namespace hlsd
