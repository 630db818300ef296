// HLSD_PATH_BLOCKED: blocked right-looking Cholesky for large N (N a multiple of 128, N >= 256) in the reference's
// EXACT arithmetic - bit-identical to the csim of the right-looking designs (cholesky_v1.3 / v3.2 / v4.0,
// Cholesky/cholesky_v1.3/model4x4/chol.cpp:25-63), at 10-25x the speed of the one-CTA-per-matrix global-memory kernel.
//
// Why blocking keeps the bits: in the reference element (i, j), i >= j, receives  a_ij = a_ij - l_ik * l_jk  (one
// rounded multiply, one rounded subtract) for k = 0 .. j-1 in ascending order, then a square root (i = j) or a
// division by l_jj.  Any schedule that applies those updates to each element in ascending k reproduces the result;
// a right-looking sweep over 128-column panels does (panel q's updates reach every trailing element before panel
// q+1's), and inside a panel so do the three kernels below:
//   1. blk_diag_kernel<DIAG_EXACT>  (blocked.cu)  L11 = chol(X[q][q]), the bit-identical 128 x 128 kernel
//   2. xchol_panel_kernel   rows below the diagonal block: for k in the panel, ascending:
//                           x_ik = (x_ik - sum_{k' < k in panel, ascending} x_ik' l_kk') / l_kk
//                           64 rows per CTA; by 32-column sub-blocks: 2 x 4 register tiles for the part that comes
//                           from earlier sub-blocks, then one thread per row through the 32 x 32 triangle
//   3. xchol_update_kernel  trailing matrix: c_ij = c_ij - l_ik * l_jk for the panel's 128 k in ascending order,
//                           64 x 64 tiles, 4 x 4 per thread, operands transposed in shared memory so that a step is
//                           two 16-byte loads and sixteen multiply-subtract pairs
// The operand A is never copied: panel 0 reads its inputs from A and writes to L; later panels work in place in L.
//
// Blocked exact LU (lu_blocked_exact in blocked.cu, 256 <= N <= 1024): the pivoted panel kernel of the tensor path
// in exact arithmetic (lu_panel_kernel<true>), its exchange / un-exchange kernels, and two kernels from this file:
//   xchol_panel_kernel<1>   U12 = L11^-1 A12 by forward substitution in the reference's order
//                           (u_kj = a_kj - sum_{k' < k, ascending} l_kk' u_k'j; the same recurrence as the Cholesky
//                           panel solve with rows and columns swapped and a unit diagonal: the tile is transposed
//                           on its way into shared memory)
//   xchol_update_kernel<1>  A22[i][j] = A22[i][j] - l_ik * u_kj, k ascending, all 64 x 64 tiles
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

constexpr int XB = 128;  // panel width
constexpr int XR = 64;   // rows per CTA of the panel kernel / tile edge of the update kernel

struct XPanelSmem {
    static constexpr int LDX = XB + 4;                 // row stride (floats) of both tiles: 16-byte aligned rows
    static constexpr int kX = XR * LDX, kL = XB * LDX;  // X tile (XR rows of the panel), L11
    static constexpr int kBytes = (kX + kL) * 4;
};

// MODE 0 (Cholesky): grid (row tiles below the diagonal block, batch); X tile = XR rows x 128 panel columns of S
//                    (= A for panel 0, L afterwards), result to L.
// MODE 1 (LU, U12):  grid (64-column tiles right of the panel, batch); the tile is 128 panel ROWS x XR columns of
//                    the working matrix S (= the U buffer), kept TRANSPOSED in shared memory (Xs[j][k]); unit
//                    diagonal: no division; result back into S's buffer through `X`.
// MODE 2 (cholesky_v2.2): grid and tile as MODE 0, S = X = the working matrix W; `L` holds the multipliers m of the
//                    diagonal block (they play L11's part, WITHOUT a division: w_c = x_c - sum_{m<c} w_m m_cm); the
//                    un-scaled values w go back to the working matrix and m = w / d (d = the diagonal of W) to `L`.
// 256 threads
template <int MODE>
__global__ void __launch_bounds__(256, 2)
xchol_panel_kernel(const float* __restrict__ S, float* __restrict__ X, const float* __restrict__ L, int n, int o,
                   float* __restrict__ Mout = nullptr) {
    constexpr int LDX = XPanelSmem::LDX;
    extern __shared__ __align__(16) float xp_smem[];
    float* Xs = xp_smem;                    // [XR][LDX]
    float* Ls = xp_smem + XPanelSmem::kX;   // [XB][LDX]: L11 (lower triangle incl. diagonal)
    const int tid = threadIdx.x;
    const long mat = (long)blockIdx.y * n * n;
    const int row0 = o + XB + blockIdx.x * XR;  // MODE 0: first row of this CTA; MODE 1: first column
    const float* xsrc = MODE != 1 ? S + mat + (long)row0 * n + o : S + mat + (long)o * n + row0;
    const float* lsrc = L + mat + (long)o * n + o;     // L11, written by the diagonal-block / panel kernel
    if constexpr (MODE != 1) {
        for (int e = tid; e < XR * (XB / 4); e += 256) {
            const int r = e >> 5, c4 = e & 31;
            *reinterpret_cast<float4*>(Xs + r * LDX + 4 * c4) = *(reinterpret_cast<const float4*>(xsrc + (long)r * n) + c4);
        }
    } else {
        // transposing load: consecutive threads take consecutive panel rows k, so the scattered stores are conflict-free
        for (int e = tid; e < XB * (XR / 4); e += 256) {
            const int k = e & (XB - 1), j4 = e >> 7;
            const float4 v = *(reinterpret_cast<const float4*>(xsrc + (long)k * n) + j4);
            Xs[(4 * j4) * LDX + k] = v.x, Xs[(4 * j4 + 1) * LDX + k] = v.y;
            Xs[(4 * j4 + 2) * LDX + k] = v.z, Xs[(4 * j4 + 3) * LDX + k] = v.w;
        }
    }
    for (int e = tid; e < XB * (XB / 4); e += 256) {
        const int r = e >> 5, c4 = e & 31;
        *reinterpret_cast<float4*>(Ls + r * LDX + 4 * c4) = *(reinterpret_cast<const float4*>(lsrc + (long)r * n) + c4);
    }
    __syncthreads();
    for (int s = 0; s < 4; s++) {
        const int c0 = 32 * s;
        if (s > 0) {
            // (a) columns c0 .. c0+31 receive the updates of the panel's earlier sub-blocks, k = 0 .. c0-1 ascending:
            //     thread -> rows (r, r+1) x columns (c .. c+3); per four k: 6 float4 loads, 32 multiply-subtract pairs
            const int r = 2 * (tid >> 3), c = c0 + 4 * (tid & 7);
            float x[2][4];
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const float4 v = *reinterpret_cast<const float4*>(Xs + (r + a) * LDX + c);
                x[a][0] = v.x, x[a][1] = v.y, x[a][2] = v.z, x[a][3] = v.w;
            }
            for (int k4 = 0; k4 < c0 / 4; k4++) {
                float xa[2][4], lb[4][4];
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const float4 v = *reinterpret_cast<const float4*>(Xs + (r + a) * LDX + 4 * k4);
                    xa[a][0] = v.x, xa[a][1] = v.y, xa[a][2] = v.z, xa[a][3] = v.w;
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const float4 v = *reinterpret_cast<const float4*>(Ls + (c + b) * LDX + 4 * k4);
                    lb[b][0] = v.x, lb[b][1] = v.y, lb[b][2] = v.z, lb[b][3] = v.w;
                }
#pragma unroll
                for (int kk = 0; kk < 4; kk++)  // ascending k
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) x[a][b] = xmsub(x[a][b], xa[a][kk], lb[b][kk]);
            }
            // (columns < c0 are final and every element of columns >= c0 belongs to one thread: no hazard)
#pragma unroll
            for (int a = 0; a < 2; a++)
                *reinterpret_cast<float4*>(Xs + (r + a) * LDX + c) = make_float4(x[a][0], x[a][1], x[a][2], x[a][3]);
            __syncthreads();
        }
        // (b) the 32 x 32 triangle of the sub-block, one thread per row:
        //     x_c = (x_c - sum_{m < c, ascending} x_m * l_{c0+c, c0+m}) / l_{c0+c, c0+c}
        if (tid < XR) {
            float* xrow = Xs + tid * LDX + c0;
            float x[32];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 v = reinterpret_cast<const float4*>(xrow)[q];
                x[4 * q] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
            }
            static_for<0, 32>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const float4* lrow = reinterpret_cast<const float4*>(Ls + (c0 + c) * LDX + c0);
                float acc = x[c];
                static_for<0, c / 4 + 1>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    const float4 lv = lrow[q];
                    if constexpr (4 * q < c) acc = xmsub(acc, x[4 * q], lv.x);
                    if constexpr (4 * q + 1 < c) acc = xmsub(acc, x[4 * q + 1], lv.y);
                    if constexpr (4 * q + 2 < c) acc = xmsub(acc, x[4 * q + 2], lv.z);
                    if constexpr (4 * q + 3 < c) acc = xmsub(acc, x[4 * q + 3], lv.w);
                    if constexpr (4 * q == (c & ~3)) {  // the piece that holds the diagonal element
                        const float d = (c & 3) == 0 ? lv.x : ((c & 3) == 1 ? lv.y : ((c & 3) == 2 ? lv.z : lv.w));
                        x[c] = MODE == 0 ? xdiv(acc, d) : acc;  // (LU: unit diagonal)
                    }
                });
            });
#pragma unroll
            for (int q = 0; q < 8; q++)
                reinterpret_cast<float4*>(xrow)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        }
        __syncthreads();
    }
    if constexpr (MODE != 1) {
        float* dst = X + mat + (long)row0 * n + o;
        for (int e = tid; e < XR * (XB / 4); e += 256) {
            const int r = e >> 5, c4 = e & 31;
            *(reinterpret_cast<float4*>(dst + (long)r * n) + c4) = *reinterpret_cast<const float4*>(Xs + r * LDX + 4 * c4);
        }
        if constexpr (MODE == 2) {
            // m = w / d, column by column of the panel (IEEE division; d = the stage's un-rooted pivot)
            float* ds = Ls;  // L11 is no longer needed: its first row holds the 128 pivots
            __syncthreads();
            if (tid < XB) ds[tid] = S[mat + (long)(o + tid) * n + o + tid];
            __syncthreads();
            float* mdst = Mout + mat + (long)row0 * n + o;
            for (int e = tid; e < XR * (XB / 4); e += 256) {
                const int r = e >> 5, c4 = e & 31;
                const float4 w = *reinterpret_cast<const float4*>(Xs + r * LDX + 4 * c4);
                const float4 d = *reinterpret_cast<const float4*>(ds + 4 * c4);
                *(reinterpret_cast<float4*>(mdst + (long)r * n) + c4) = make_float4(xdiv(w.x, d.x), xdiv(w.y, d.y), xdiv(w.z, d.z), xdiv(w.w, d.w));
            }
        }
    } else {
        float* dst = X + mat + (long)o * n + row0;
        for (int e = tid; e < XB * (XR / 4); e += 256) {
            const int k = e & (XB - 1), j4 = e >> 7;
            *(reinterpret_cast<float4*>(dst + (long)k * n) + j4) =
                make_float4(Xs[(4 * j4) * LDX + k], Xs[(4 * j4 + 1) * LDX + k], Xs[(4 * j4 + 2) * LDX + k], Xs[(4 * j4 + 3) * LDX + k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Exact Cholesky of an M x M block (M = 64 or 128) held in shared memory, one CTA per matrix, blocked by 32 columns:
//   for j = 0 .. M/32 - 1:
//     (a) block column j receives the updates of the columns left of it, k ascending, 2 x 2 register tiles
//     (b) ONE warp factors the 32 x 32 diagonal block: lane = row, the row in registers, pivot and column stream by
//         shuffle, IEEE square root and division, separately rounded multiply-subtract
//     (c) rows below: one thread per row through the 32 x 32 triangle (forward substitution, IEEE division)
// Per element the operations and their order are the reference's (updates in ascending k, then root / division), so
// the result is bit-identical.  Compared with the register kernels (chol_rows_kernel<64>: every step shuffles all 64
// columns through the warp; blk_diag_kernel<EXACT>: one block-wide barrier per column) the shuffle / barrier work is
// confined to the 32 x 32 diagonal blocks and the rest runs from shared memory with small loops that stay in the
// instruction cache; 17 KB (M = 64) / 68 KB (M = 128) of shared memory, so many matrices share an SM and hide each
// other's serial diagonal-block phases.  Serves N = 64 and N = 128 and, with a row stride, the diagonal blocks of
// the blocked exact path.  FUSED: HLSD_FLAG_FUSED (a - b*c contracted).
// ------------------------------------------------------------------------------------------------
template <bool FUSED>
__device__ __forceinline__ void xchol32(float* T, int ldt, int c0, int lane) {
    float d[32];
    {
        const float4* row4 = reinterpret_cast<const float4*>(T + (c0 + lane) * ldt + c0);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 v = row4[q];
            d[4 * q] = v.x, d[4 * q + 1] = v.y, d[4 * q + 2] = v.z, d[4 * q + 3] = v.w;
        }
    }
    static_for<0, 32>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        // lanes above row k hold scratch in d[k] (possibly zero or negative): they work on the pivot itself, so that no
        // lane leaves the fast paths of the square root and the division (their results are never stored)
        const float pk = __shfl_sync(0xffffffffu, d[k], k);
        const float dk = xsqrt(pk);
        const float num = lane > k ? d[k] : dk;
        const float lk = lane == k ? dk : xdiv(num, make_divisor(dk));
        d[k] = lk;
#pragma unroll
        for (int c = k + 1; c < 32; c++) d[c] = msub<FUSED>(d[c], lk, __shfl_sync(0xffffffffu, lk, c));
    });
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (c <= lane) T[(c0 + lane) * ldt + c0 + c] = d[c];
}

template <int M, int NT, bool FUSED>
__global__ void __launch_bounds__(NT)
xchol_tile_kernel(const float* __restrict__ S, float* __restrict__ L, int n, int o, long batch_stride) {
    constexpr int LDT = M + 4;
    extern __shared__ __align__(16) float xt_smem[];
    float* T = xt_smem;  // [M][LDT]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* sblk = S + (long)blockIdx.x * batch_stride + (long)o * n + o;
    float* lblk = L + (long)blockIdx.x * batch_stride + (long)o * n + o;
    for (int e = tid; e < M * (M / 4); e += NT) {
        const int row = e / (M / 4), c4 = e % (M / 4);
        *reinterpret_cast<float4*>(T + row * LDT + 4 * c4) = *(reinterpret_cast<const float4*>(sblk + (long)row * n) + c4);
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < M / 32; j++) {
        const int c0 = 32 * j;
        if (j > 0) {
            // (a) rows r0, r0+1 x columns ca, ca+16 of block column j: every element takes k = 0 .. c0-1 in ascending order
            const int tiles = ((M - c0) / 2) * 16;
            for (int t = tid; t < tiles; t += NT) {
                const int cp = t & 15, rp = t >> 4;
                const int r0 = c0 + 2 * rp, ca = c0 + cp, cb = ca + 16;
                if (r0 + 1 < ca) continue;  // entirely above the diagonal
                const float4* A0 = reinterpret_cast<const float4*>(T + r0 * LDT);
                const float4* A1 = reinterpret_cast<const float4*>(T + (r0 + 1) * LDT);
                const float4* B0 = reinterpret_cast<const float4*>(T + ca * LDT);
                const float4* B1 = reinterpret_cast<const float4*>(T + cb * LDT);
                float s00 = T[r0 * LDT + ca], s01 = T[r0 * LDT + cb], s10 = T[(r0 + 1) * LDT + ca], s11 = T[(r0 + 1) * LDT + cb];
                for (int k4 = 0; k4 < c0 / 4; k4++) {
                    const float4 a0 = A0[k4], a1 = A1[k4], b0 = B0[k4], b1 = B1[k4];
                    s00 = msub<FUSED>(s00, a0.x, b0.x), s01 = msub<FUSED>(s01, a0.x, b1.x), s10 = msub<FUSED>(s10, a1.x, b0.x), s11 = msub<FUSED>(s11, a1.x, b1.x);
                    s00 = msub<FUSED>(s00, a0.y, b0.y), s01 = msub<FUSED>(s01, a0.y, b1.y), s10 = msub<FUSED>(s10, a1.y, b0.y), s11 = msub<FUSED>(s11, a1.y, b1.y);
                    s00 = msub<FUSED>(s00, a0.z, b0.z), s01 = msub<FUSED>(s01, a0.z, b1.z), s10 = msub<FUSED>(s10, a1.z, b0.z), s11 = msub<FUSED>(s11, a1.z, b1.z);
                    s00 = msub<FUSED>(s00, a0.w, b0.w), s01 = msub<FUSED>(s01, a0.w, b1.w), s10 = msub<FUSED>(s10, a1.w, b0.w), s11 = msub<FUSED>(s11, a1.w, b1.w);
                }
                if (r0 >= ca) T[r0 * LDT + ca] = s00;
                if (r0 >= cb) T[r0 * LDT + cb] = s01;
                if (r0 + 1 >= ca) T[(r0 + 1) * LDT + ca] = s10;
                if (r0 + 1 >= cb) T[(r0 + 1) * LDT + cb] = s11;
            }
            __syncthreads();
        }
        // (b) the diagonal block, inside one warp
        if (warp == 0) xchol32<FUSED>(T, LDT, c0, lane);
        __syncthreads();
        // (c) rows below, one thread per row: x_c = (x_c - sum_{m < c, ascending} x_m l_cm) / l_cc
        if (tid < M - c0 - 32) {
            float* xrow = T + (c0 + 32 + tid) * LDT + c0;
            float x[32];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 v = reinterpret_cast<const float4*>(xrow)[q];
                x[4 * q] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
            }
            static_for<0, 32>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const float4* lrow = reinterpret_cast<const float4*>(T + (c0 + c) * LDT + c0);
                float acc = x[c];
                static_for<0, c / 4 + 1>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    const float4 lv = lrow[q];
                    if constexpr (4 * q < c) acc = msub<FUSED>(acc, x[4 * q], lv.x);
                    if constexpr (4 * q + 1 < c) acc = msub<FUSED>(acc, x[4 * q + 1], lv.y);
                    if constexpr (4 * q + 2 < c) acc = msub<FUSED>(acc, x[4 * q + 2], lv.z);
                    if constexpr (4 * q + 3 < c) acc = msub<FUSED>(acc, x[4 * q + 3], lv.w);
                    if constexpr (4 * q == (c & ~3)) {
                        const float dd = (c & 3) == 0 ? lv.x : ((c & 3) == 1 ? lv.y : ((c & 3) == 2 ? lv.z : lv.w));
                        x[c] = xdiv(acc, dd);
                    }
                });
            });
#pragma unroll
            for (int q = 0; q < 8; q++)
                reinterpret_cast<float4*>(xrow)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        }
        __syncthreads();
    }
    for (int e = tid; e < M * (M / 4); e += NT) {
        const int row = e / (M / 4), c4 = e % (M / 4);
        float4 v = *reinterpret_cast<const float4*>(T + row * LDT + 4 * c4);
        if (4 * c4 > row) v.x = 0.0f;
        if (4 * c4 + 1 > row) v.y = 0.0f;
        if (4 * c4 + 2 > row) v.z = 0.0f;
        if (4 * c4 + 3 > row) v.w = 0.0f;
        *(reinterpret_cast<float4*>(lblk + (long)row * n) + c4) = v;
    }
}

// M x M diagonal block at (o, o) of every matrix of the batch (row stride n, matrix stride n * n); n == M, o == 0: a
// batch of M x M matrices
template <int M, int NT, bool FUSED>
static cudaError_t launch_xchol_tile(const float* S, float* L, int n, int o, long batch, cudaStream_t st) {
    constexpr size_t smem = (size_t)M * (M + 4) * 4;
    auto kern = xchol_tile_kernel<M, NT, FUSED>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    for (long b0 = 0; b0 < batch; b0 += 1 << 20) {
        const long cnt = std::min<long>(batch - b0, 1 << 20);
        kern<<<(unsigned)cnt, NT, smem, st>>>(S + b0 * (long)n * n, L + b0 * (long)n * n, n, o, (long)n * n);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t cholesky_tile_exact(const float* S, float* L, int m, int n, int o, long batch, int fused, cudaStream_t st) {
    if (m == 64) {
        // measured (batch 98304): 128 threads per matrix 49.9 M fact/s, 64 threads 41.9 M, register kernel 36.2 M
        if (option(HLSD_OPT_CHOL_TILE) == 2)
            return fused ? launch_xchol_tile<64, 64, true>(S, L, n, o, batch, st) : launch_xchol_tile<64, 64, false>(S, L, n, o, batch, st);
        return fused ? launch_xchol_tile<64, 128, true>(S, L, n, o, batch, st) : launch_xchol_tile<64, 128, false>(S, L, n, o, batch, st);
    }
    if (m == 128) return fused ? launch_xchol_tile<128, 256, true>(S, L, n, o, batch, st) : launch_xchol_tile<128, 256, false>(S, L, n, o, batch, st);
    return cudaErrorNotSupported;
}

struct XUpdSmem {
    static constexpr int LDK = XR + 4;                 // row stride of the transposed operand tiles [k][row]
    static constexpr int kOp = XB * LDK;
    static constexpr int kBytes = 2 * kOp * 4;
};

// LU == 0 (Cholesky): grid (lower-triangular 64 x 64 tiles of the trailing matrix, batch); C: Csrc -> Cdst (= L),
//                      both operands are rows of L.
// LU == 1:             grid (all tiles, batch); C in place in the working matrix (Csrc = Cdst = the U buffer); A operand
//                      = rows of L (multipliers, final row order), B operand = rows o..o+127 of the working matrix (U12),
//                      which are already [k][j].
// 256 threads; thread -> 4 x 4 outputs
// LU == 2 (cholesky_v2.2): grid and C as LU == 0 but in place in the working matrix W (Csrc = Cdst = W); A operand = rows
//                      of W (the un-scaled column values w), B operand = rows of `L` (the multipliers m).
template <int LU>
__global__ void __launch_bounds__(256, 3)
xchol_update_kernel(const float* __restrict__ Csrc, float* __restrict__ Cdst, const float* __restrict__ L, int n, int o) {
    constexpr int LDK = XUpdSmem::LDK;
    extern __shared__ __align__(16) float xu_smem[];
    float* As = xu_smem;                  // [XB k][LDK]: As[k][i] = L[ri0 + i][o + k]
    float* Bs = xu_smem + XUpdSmem::kOp;  // [XB k][LDK]: Bs[k][j] = L[rj0 + j][o + k]  (LU: U[o + k][rj0 + j])
    const int tid = threadIdx.x;
    int ti, tj;
    if constexpr (LU == 1) {
        const int m = (n - o - XB) / XR;
        ti = blockIdx.x / m, tj = blockIdx.x % m;
    } else {
        // unrank blockIdx.x into (ti, tj), ti >= tj
        int rr = blockIdx.x;
        ti = (int)((sqrtf(8.0f * rr + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= rr) ti++;
        while (ti * (ti + 1) / 2 > rr) ti--;
        tj = rr - ti * (ti + 1) / 2;
    }
    const long mat = (long)blockIdx.y * n * n;
    const int base = o + XB;  // first row / column of the trailing matrix
    const int ri0 = base + ti * XR, rj0 = base + tj * XR;
    const float* Lm = L + mat;
    // transposing loads: thread reads 16 bytes along k of one row and scatters them to four k rows of the tile;
    // consecutive threads take consecutive rows, so the scattered stores are conflict-free
    for (int e = tid; e < XR * (XB / 4); e += 256) {
        const int i = e & (XR - 1), k4 = e >> 6;
        const float* Am = LU == 2 ? Csrc + mat : Lm;  // cholesky_v2.2: the A operand is the working matrix itself
        const float4 va = *(reinterpret_cast<const float4*>(Am + (long)(ri0 + i) * n + o) + k4);
        As[(4 * k4) * LDK + i] = va.x, As[(4 * k4 + 1) * LDK + i] = va.y;
        As[(4 * k4 + 2) * LDK + i] = va.z, As[(4 * k4 + 3) * LDK + i] = va.w;
        if (LU == 2 || (!LU && ti != tj)) {
            const float4 vb = *(reinterpret_cast<const float4*>(Lm + (long)(rj0 + i) * n + o) + k4);
            Bs[(4 * k4) * LDK + i] = vb.x, Bs[(4 * k4 + 1) * LDK + i] = vb.y;
            Bs[(4 * k4 + 2) * LDK + i] = vb.z, Bs[(4 * k4 + 3) * LDK + i] = vb.w;
        }
    }
    if constexpr (LU == 1) {
        const float* Um = Csrc + mat + (long)o * n + rj0;  // rows o..o+127 of the working matrix, columns of this tile
        for (int e = tid; e < XB * (XR / 4); e += 256) {
            const int k = e >> 4, j4 = e & 15;
            *reinterpret_cast<float4*>(Bs + k * LDK + 4 * j4) = *(reinterpret_cast<const float4*>(Um + (long)k * n) + j4);
        }
    }
    const float* Bt = (LU || ti != tj) ? Bs : As;  // (LU == 2: the operands differ even on the diagonal tiles)
    const int i0 = 4 * (tid >> 4), j0 = 4 * (tid & 15);
    float c[4][4];
    const float* csrc = Csrc + mat + (long)(ri0 + i0) * n + rj0 + j0;  // Csrc = A for panel 0, L afterwards
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const float4 v = *reinterpret_cast<const float4*>(csrc + (long)a * n);
        c[a][0] = v.x, c[a][1] = v.y, c[a][2] = v.z, c[a][3] = v.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < XB; k++) {  // ascending k: the reference's order for every element
        const float4 av = *reinterpret_cast<const float4*>(As + k * LDK + i0);
        const float4 bv = *reinterpret_cast<const float4*>(Bt + k * LDK + j0);
        // two columns per instruction (FMUL2 / FFMA2, common.cuh): (c[a][0], c[a][1]) -= a * (b0, b1), each lane rounded
        // twice like the scalar pair
        const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            msub_pair<false>(c[a][0], c[a][1], a4[a], bv.x, bv.y);
            msub_pair<false>(c[a][2], c[a][3], a4[a], bv.z, bv.w);
        }
    }
    float* cdst = Cdst + mat + (long)(ri0 + i0) * n + rj0 + j0;
#pragma unroll
    for (int a = 0; a < 4; a++) *reinterpret_cast<float4*>(cdst + (long)a * n) = make_float4(c[a][0], c[a][1], c[a][2], c[a][3]);
}

// blocked.cu
cudaError_t chol_diag_exact(const float* S, float* L, int n, int o, long batch, cudaStream_t st);
cudaError_t chol_zero_upper_blocks(float* L, int n, long batch, cudaStream_t st);

cudaError_t cholesky_blocked_exact(const float* A, float* L, int n, long batch, cudaStream_t st) {
    if (n % XB != 0 || n < 2 * XB) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    if (batch > 65535) return cudaErrorNotSupported;
    cudaError_t e;
    if ((e = ensure_dynamic_smem(xchol_panel_kernel<0>, XPanelSmem::kBytes)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(xchol_update_kernel<0>, XUpdSmem::kBytes)) != cudaSuccess) return e;
    const int panels = n / XB;
    const unsigned nb = (unsigned)batch;
    if ((e = chol_zero_upper_blocks(L, n, batch, st)) != cudaSuccess) return e;
    for (int q = 0; q < panels; q++) {
        const int o = q * XB;
        const float* src = (q == 0) ? A : L;  // where the not yet factored part currently lives
        {
            ProfScope ps(HLSD_PROF_DIAG, st);
            e = option(HLSD_OPT_CHOL_TILE) == 1 ? chol_diag_exact(src, L, n, o, batch, st)
                                                : cholesky_tile_exact(src, L, XB, n, o, batch, 0, st);
        }
        if (e != cudaSuccess) return e;
        const int below = n - o - XB;
        if (below > 0) {
            {
                ProfScope ps(HLSD_PROF_PANEL, st);
                xchol_panel_kernel<0><<<dim3(below / XR, nb), 256, XPanelSmem::kBytes, st>>>(src, L, L, n, o);
                count_launch();
            }
            {
                ProfScope ps(HLSD_PROF_UPDATE, st);
                const int m = below / XR;
                xchol_update_kernel<0><<<dim3(m * (m + 1) / 2, nb), 256, XUpdSmem::kBytes, st>>>(src, L, L, n, o);
                count_launch();
            }
        }
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// cholesky_v2.2 ("revised": root-free elimination, roots at the end; Cholesky/cholesky_v2.2/model4x4/chol.cpp:28-207)
// for N >= 256, blocked like the right-looking designs above.  Per element
// (j, i), j >= i, the reference applies  w_ji -= w_jk * m_ik  for k ascending, with m_ik = w_ik / w_kk taken at stage k,
// and finally l_jk = m_jk * sqrt(w_kk).  The working matrix W (a workspace copy of A) keeps the un-scaled values w, the L
// buffer the multipliers m; per 128-column panel: the diagonal block (x22_diag_kernel), the rows below it
// (xchol_panel_kernel<2>: the panel recurrence without a division, then m = w / d), the trailing update
// (xchol_update_kernel<2>: A operand = w, B operand = m); at the end every column is scaled by its root.
// ------------------------------------------------------------------------------------------------
// Diagonal block: one CTA per matrix, the 128 x 128 block in shared memory - w in the lower triangle (with the diagonal),
// the multipliers of column k in row k of the upper triangle (m_ik at T[k][i]).  256 threads, two barriers per column.
__global__ void __launch_bounds__(256)
x22_diag_kernel(float* __restrict__ W, float* __restrict__ L, int n, int o) {
    constexpr int LDT = XB + 1;
    extern __shared__ __align__(16) float x22_tile[];  // [XB][LDT]
    float* T = x22_tile;
    const int tid = threadIdx.x;
    float* wblk = W + (long)blockIdx.x * n * n + (long)o * n + o;
    float* lblk = L + (long)blockIdx.x * n * n + (long)o * n + o;
    for (int e = tid; e < XB * XB; e += 256) {
        const int r = e >> 7, c = e & (XB - 1);
        T[r * LDT + c] = wblk[(long)r * n + c];
    }
    __syncthreads();
    for (int k = 0; k < XB - 1; k++) {
        const float d = T[k * LDT + k];
        for (int i = k + 1 + tid; i < XB; i += 256) T[k * LDT + i] = xdiv(T[i * LDT + k], d);  // m_ik
        __syncthreads();
        // rows j = k+1 .. 127, eight threads per row striding over the columns i = k+1 .. j of the lower triangle
        for (int j = k + 1 + (tid >> 3); j < XB; j += 32) {
            const float wjk = T[j * LDT + k];
            for (int i = k + 1 + (tid & 7); i <= j; i += 8) T[j * LDT + i] = xmsub(T[j * LDT + i], wjk, T[k * LDT + i]);
        }
        __syncthreads();
    }
    for (int e = tid; e < XB * XB; e += 256) {
        const int r = e >> 7, c = e & (XB - 1);
        if (c <= r) wblk[(long)r * n + c] = T[r * LDT + c];
        lblk[(long)r * n + c] = c < r ? T[c * LDT + r] : 0.0f;  // m below the diagonal; the diagonal follows with the roots
    }
}

// L[j][k] = m_jk * sqrt(w_kk) below the diagonal, sqrt(w_kk) on it.  grid: (n / 128 column blocks, batch), 256 threads
__global__ void __launch_bounds__(256)
x22_finish_kernel(const float* __restrict__ W, float* __restrict__ L, int n) {
    __shared__ float root[XB];
    const int c0 = blockIdx.x * XB, tid = threadIdx.x;
    const float* w = W + (long)blockIdx.y * n * n;
    float* l = L + (long)blockIdx.y * n * n;
    if (tid < XB) root[tid] = xsqrt(w[(long)(c0 + tid) * n + c0 + tid]);
    __syncthreads();
    for (long e = tid; e < (long)(n - c0) * (XB / 4); e += 256) {
        const int r = c0 + (int)(e >> 5), c4 = (int)(e & 31);
        float4* p = reinterpret_cast<float4*>(l + (long)r * n + c0) + c4;
        float4 v = *p;
        float out[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = c0 + 4 * c4 + q;
            out[q] = c < r ? xmul(out[q], root[4 * c4 + q]) : (c == r ? root[4 * c4 + q] : 0.0f);
        }
        *p = make_float4(out[0], out[1], out[2], out[3]);
    }
}

cudaError_t cholesky_blocked_exact_v22(const float* A, float* L, int n, long batch, cudaStream_t st) {
    if (n % XB != 0 || n < 2 * XB) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = ensure_dynamic_smem(xchol_panel_kernel<2>, XPanelSmem::kBytes)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(xchol_update_kernel<2>, XUpdSmem::kBytes)) != cudaSuccess) return e;
    const size_t diag_smem = (size_t)XB * (XB + 1) * sizeof(float);
    if ((e = ensure_dynamic_smem(x22_diag_kernel, diag_smem)) != cudaSuccess) return e;
    const long nn = (long)n * n;
    const int panels = n / XB;
    // the working matrix is a workspace copy of A; large batches go in slices that bound it to 16 GB
    const long slice = std::max<long>(1, std::min<long>(std::min<long>(batch, 32768), (long)(((size_t)16 << 30) / (sizeof(float) * nn))));
    float* W = nullptr;
    if ((e = ws_alloc(reinterpret_cast<void**>(&W), sizeof(float) * (size_t)slice * nn, st)) != cudaSuccess) return e;
    for (long b0 = 0; b0 < batch && e == cudaSuccess; b0 += slice) {
        const long cnt = std::min(slice, batch - b0);
        const unsigned nb = (unsigned)cnt;
        float* Ls = L + b0 * nn;
        {
            ProfScope ps(HLSD_PROF_OTHER, st);
            e = cudaMemcpyAsync(W, A + b0 * nn, sizeof(float) * (size_t)cnt * nn, cudaMemcpyDeviceToDevice, st);
        }
        if (e != cudaSuccess) break;
        if ((e = chol_zero_upper_blocks(Ls, n, cnt, st)) != cudaSuccess) break;
        for (int q = 0; q < panels; q++) {
            const int o = q * XB;
            {
                ProfScope ps(HLSD_PROF_DIAG, st);
                x22_diag_kernel<<<nb, 256, diag_smem, st>>>(W, Ls, n, o);
                count_launch();
            }
            const int below = n - o - XB;
            if (below > 0) {
                {
                    ProfScope ps(HLSD_PROF_PANEL, st);
                    xchol_panel_kernel<2><<<dim3(below / XR, nb), 256, XPanelSmem::kBytes, st>>>(W, W, Ls, n, o, Ls);
                    count_launch();
                }
                {
                    ProfScope ps(HLSD_PROF_UPDATE, st);
                    const int m = below / XR;
                    xchol_update_kernel<2><<<dim3(m * (m + 1) / 2, nb), 256, XUpdSmem::kBytes, st>>>(W, W, Ls, n, o);
                    count_launch();
                }
            }
        }
        {
            ProfScope ps(HLSD_PROF_OTHER, st);
            x22_finish_kernel<<<dim3(panels, nb), 256, 0, st>>>(W, Ls, n);
            count_launch();
        }
        e = cudaGetLastError();
    }
    ws_free(W, st);
    return e;
}

// the two exact kernels of the blocked LU (called from lu_blocked_exact, blocked.cu): U12 and the trailing update of
// the panel at offset o; Wk = working matrix (the U buffer), L = multipliers in final row order
cudaError_t xlu_u12_and_update(float* Wk, const float* L, int n, int o, long batch, cudaStream_t st) {
    cudaError_t e;
    if ((e = ensure_dynamic_smem(xchol_panel_kernel<1>, XPanelSmem::kBytes)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(xchol_update_kernel<1>, XUpdSmem::kBytes)) != cudaSuccess) return e;
    const int m = (n - o - XB) / XR;
    if (m <= 0) return cudaSuccess;
    {
        ProfScope ps(HLSD_PROF_TRSM, st);
        xchol_panel_kernel<1><<<dim3(m, (unsigned)batch), 256, XPanelSmem::kBytes, st>>>(Wk, Wk, L, n, o);
        count_launch();
    }
    {
        ProfScope ps(HLSD_PROF_UPDATE, st);
        xchol_update_kernel<1><<<dim3(m * m, (unsigned)batch), 256, XUpdSmem::kBytes, st>>>(Wk, Wk, L, n, o);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace hlsd
