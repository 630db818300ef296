// Blocked right-looking Cholesky for large N (N a multiple of 128), batched.
//
// The reference's systolic Cholesky is a right-looking elimination (Cholesky/cholesky_v1.3/
// model4x4/chol.cpp:25-63): per column k a square root, a column scaling and a rank-1 update of
// the trailing matrix.  At N = 1024 the rank-1 updates are >99 % of the work and, grouped over a
// panel of NB = 128 columns, they are a dense contraction  C -= P * P^T  (P = the finished panel
// below the diagonal block).  That part runs on the tcgen05 tensor cores; everything that is not a
// dense contraction stays on CUDA cores:
//
//   per block column q (columns o = 128 q .. o+127), all matrices of the batch at once (block column 0 reads
//   its inputs straight from A, so the operand is never copied), LEFT-looking:
//     1. blk_ll2_kernel  (tcgen05)  X[r][q] = A[r][q] - sum_{j<q} L[r][j] L[q][j]^T   (all earlier panels at once)
//     2. blk_diag_kernel (SIMT, register-tiled)   L11 = chol(X[q][q]), W = inv(L11)
//     3. blk_mma_kernel  (tcgen05)  L[r][q] = X[r][q] * W^T                            (rows below the diagonal block)
//   (A right-looking variant - one SYRK launch per panel - and two one-tile-per-CTA update kernels were measured in
//   round 1 and removed in round 2: 20 % / 8 % slower)
//
// Tensor-core arithmetic: fp32 operands are split on the fly into tf32 hi (truncation) + tf32 lo parts and each
// product is formed as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (3xTF32) with fp32 accumulation in TMEM,
// which keeps ~21 mantissa bits (a single tf32 pass keeps 10).  The result is not bit-identical to
// the csim (different summation order, split products) but agrees to fp32 round-off level; the
// exact paths (HLSD_PATH_GLOBAL) remain available at any N.
//
// blk_mma_kernel / gemm_tile_kernel: one CTA = one 128x128 output tile, D = A * B^T with A = 128 rows x 128 k,
// B = 128 rows x 128 k, both K-major fp32 in global memory.  K is consumed in four chunks of 32:
// cp.async (LDGSTS) copies the raw fp32 chunk of A and B straight into shared memory in the
// canonical K-major SWIZZLE_128B layout, double-buffered so that chunk c+1 streams in while chunk
// c is being multiplied; the raw tile itself is the tf32 "hi" operand (the tensor core reads the
// upper 19 bits of each word) and the threads only add the "lo" tile (x - hi, rounded to tf32);
// 96 KB per CTA, two CTAs per SM overlap each other's MMA and epilogue phases; one elected
// thread issues 12 tcgen05.mma.kind::tf32 (M128 N128 K8) per chunk; completion is tracked with
// tcgen05.commit on an mbarrier; the epilogue reads the accumulator with tcgen05.ld (32x32b.x32)
// and either stores it (TRSM) or subtracts it from C (blocked LU).

#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace hlsd {

constexpr int kDiagSmem = NB * (NB + 1) * 4;  // dynamic shared memory of blk_diag_kernel


// ------------------------------------------------------------------------------------------------
// Zero the blocks strictly above the block diagonal of L (nothing else ever writes them).  The
// operand A itself is never copied: panel 0 reads its inputs from A and writes to L, later panels
// work in place in L.  grid: (upper blocks, batch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
blk_zero_upper_kernel(float* __restrict__ L, int n) {
    int r = blockIdx.x;  // unrank into (bi < bj)
    int bj = (int)((sqrtf(8.0f * r + 1.0f) + 1.0f) * 0.5f);
    while (bj * (bj - 1) / 2 > r) bj--;
    while ((bj + 1) * bj / 2 <= r) bj++;
    const int bi = r - bj * (bj - 1) / 2;
    float4* base = reinterpret_cast<float4*>(L + (long)blockIdx.y * n * n + (long)bi * NB * n + bj * NB);
    const int n4 = n / 4;
    for (int e = threadIdx.x; e < NB * (NB / 4); e += 256) {
        const int row = e / (NB / 4), c4 = e % (NB / 4);
        base[(long)row * n4 + c4] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// diagonal block: L11 = chol(A11) (128 x 128, right-looking) and W = inv(L11), fused in one sweep.
// One CTA of 512 threads per matrix.  Warp w (16 warps) owns COLUMNS j = w + 16 b (b < 8), lane l owns
// ROWS i = l + 32 a (a < 4): a 4 x 8 register tile per thread, dealt cyclically so that the shrinking
// active region stays balanced.  Column j of the tile holds A[:, j] while k < j and, once column j has
// been finished (k >= j), W[:, j]: the two rank-1 updates of step k
//      A[i][j] -= l_ik * l_jk   (i > k, j > k)          W[i][j] -= l_ik * w_kj   (i > k, j <= k)
// touch disjoint columns, so one register tile serves both (W follows from forward substitution on
// the identity: W[k][:] = (e_k - sum_{j<k} l_kj W[j][:]) / l_kk).  With whole columns inside one warp:
//   * column k is scaled and published (the systolic "column stream") by ONE warp with all lanes busy;
//   * the diagonal element that the next step needs sits in that same warp: the next pivot root is
//     taken there and travels with the column through shared memory, so there is ONE barrier per step
//     (publication buffers double-buffered by the parity of k);
//   * row k of W, which every column needs as multiplier w_kj, sits in lane k % 32 of the very warp
//     that owns column j: a shuffle, no shared memory and no divergent single-lane section.
// L11 is collected in a shared-memory tile and written out row-wise at the end.
// Plain fp32 FMA arithmetic (tensor path: tolerance-checked).
// ------------------------------------------------------------------------------------------------
// MODE 0: L11 and W as described.  MODE 1 (DIAG_INV_ONLY): the block at S is an already finished UNIT lower
// triangular factor (blocked LU): only W = inv(L11) is produced (no roots, no scaling, the unfinished-column
// updates are switched off, L is not written).  MODE 2 (DIAG_EXACT): Cholesky only, in the reference's exact
// arithmetic (IEEE sqrt and division, separately rounded multiply and subtract, per-element update order
// k ascending as in cholesky_v1.3/model4x4/chol.cpp:25-63): the bit-identical 128 x 128 kernel.
enum { DIAG_CHOL_INV = 0, DIAG_INV_ONLY = 1, DIAG_EXACT = 2 };
template <int MODE, bool FUSED = false>  // FUSED (with DIAG_EXACT): the contraction HLSD_FLAG_FUSED allows
__global__ void __launch_bounds__(512, 2)
blk_diag_kernel(const float* __restrict__ S, float* __restrict__ L, float* __restrict__ W, int n, int o) {
    constexpr bool INV_ONLY = MODE == DIAG_INV_ONLY;
    constexpr bool EXACT = MODE == DIAG_EXACT;
    extern __shared__ __align__(16) float dg_tile[];  // [NB][NB + 1]
    __shared__ float ubuf[2][NB];                     // ubuf[k & 1][i] = l_ik for i > k
    __shared__ float sdiag[2];                        // sdiag[k & 1] = l_kk
    constexpr int ld = NB + 1;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float* blk = L + (long)blockIdx.x * n * n + (long)o * n + o;
    const float* sblk = S + (long)blockIdx.x * n * n + (long)o * n + o;  // S = A for panel 0, L afterwards
    for (int e = threadIdx.x; e < NB * NB; e += 512) {
        const int i = e >> 7, j = e & (NB - 1);
        dg_tile[i * ld + j] = (j <= i) ? sblk[(long)i * n + j] : 0.0f;
    }
    __syncthreads();
    float v[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) v[a][b] = dg_tile[(l + 32 * a) * ld + w + 16 * b];
    __syncthreads();  // the tile is rebuilt as L11 below
    // k = 16 * BK + kr with BK compile-time, so that the registers of column k / row k have constant indices
    static_for<0, 8>([&](auto bkc) {
        constexpr int BK = decltype(bkc)::value;  // column slot of column k
        constexpr int AK = BK >> 1;               // row slot of row k
        for (int kr = 0; kr < 16; kr++) {
            const int k = 16 * BK + kr;
            const int lk = k & 31;
            float* ub = ubuf[k & 1];
            if (w == kr) {  // this warp owns column k (all rows): root, scale, publish, re-seed with e_k
                const float d = INV_ONLY ? 1.0f : (EXACT ? xsqrt(__shfl_sync(0xffffffffu, v[AK][BK], lk))
                                                         : sqrtf(__shfl_sync(0xffffffffu, v[AK][BK], lk)));
                const float rd = INV_ONLY ? 1.0f : 1.0f / d;
                if (l == 0) sdiag[k & 1] = d;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int i = l + 32 * a;
                    const float lv = EXACT ? xdiv(v[a][BK], d) : v[a][BK] * rd;
                    if (i > k) ub[i] = lv;
                    dg_tile[i * ld + k] = (i > k) ? lv : (i == k ? d : 0.0f);
                    v[a][BK] = (i == k) ? 1.0f : 0.0f;
                }
            }
            __syncthreads();
            const float rd = (INV_ONLY || EXACT) ? 1.0f : 1.0f / sdiag[k & 1];
            (void)rd;
            // multipliers per column slot: slots b < BK hold columns j < 16 BK <= k (finished: w_kj, shuffled from
            // the row-k lane of this warp), slots b > BK hold columns j > k (l_jk from the column stream);
            // only slot BK needs a run-time test (j = w + 16 BK <= k  <=>  w <= kr)
            float u[8];
            static_for<0, 8>([&](auto bc) {
                constexpr int b = decltype(bc)::value;
                if constexpr (b < BK) {
                    if constexpr (EXACT) {
                        u[b] = 0.0f;
                    } else {
                        u[b] = __shfl_sync(0xffffffffu, v[AK][b], lk) * rd;
                        if (l == lk) v[AK][b] = u[b];  // row k of W is final
                    }
                } else if constexpr (b > BK) {
                    u[b] = INV_ONLY ? 0.0f : ub[w + 16 * b];
                } else {
                    if (w <= kr) {
                        if constexpr (EXACT) {
                            u[b] = 0.0f;
                        } else {
                            u[b] = __shfl_sync(0xffffffffu, v[AK][b], lk) * rd;
                            if (l == lk) v[AK][b] = u[b];
                        }
                    } else {
                        u[b] = INV_ONLY ? 0.0f : ub[w + 16 * b];
                    }
                }
            });
            // row slot AK may still hold rows <= k (run-time test); slots above it hold rows > k only
            {
                const int i = l + 32 * AK;
                const float lr = (i > k) ? ub[i] : 0.0f;
#pragma unroll
                for (int b = EXACT ? BK : 0; b <= 2 * AK + 1; b++) {
                    if constexpr (EXACT) {
                        if (i > k) v[AK][b] = msub<FUSED>(v[AK][b], lr, u[b]);  // rows <= k are not touched (0 * inf would be NaN)
                    } else {
                        v[AK][b] = fmaf(-lr, u[b], v[AK][b]);
                    }
                }
            }
            static_for<AK + 1, 4>([&](auto ac) {
                constexpr int a = decltype(ac)::value;
                const float lr = ub[l + 32 * a];
                // columns j = w + 16 b > i are never needed: i <= 32 a + 31, so b <= 2 a + 1 suffices
#pragma unroll
                for (int b = EXACT ? BK : 0; b <= 2 * a + 1; b++)
                    v[a][b] = EXACT ? msub<FUSED>(v[a][b], lr, u[b]) : fmaf(-lr, u[b], v[a][b]);
            });
        }
    });
    __syncthreads();
    // L11 rows (zeros above the diagonal), then W through the same tile
    if (!INV_ONLY)
        for (int e = threadIdx.x; e < NB * NB; e += 512) {
            const int i = e >> 7, j = e & (NB - 1);
            blk[(long)i * n + j] = (j <= i) ? dg_tile[i * ld + j] : 0.0f;
        }
    if (EXACT) return;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) dg_tile[(l + 32 * a) * ld + w + 16 * b] = v[a][b];
    __syncthreads();
    float* wout = W + (long)blockIdx.x * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += 512) {
        const int i = e >> 7, j = e & (NB - 1);
        wout[e] = (j <= i) ? dg_tile[i * ld + j] : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// Diagonal block, blocked (round 2).  ncu on blk_diag_kernel above: one barrier per column and a 4 x 8 register tile
// per thread mean ~2000 warp instructions per column for ~256 warp FMAs of useful work (issue slots 55 % busy, 13 %
// of them FMAs).  Here the 128 x 128 block is factored by 32-column sub-blocks, so that the per-column cost is paid
// on 32 x 32 blocks inside ONE warp (rows in registers, shuffles, no barrier) and everything else is a dot product
// over contiguous k with 2 x 2 register tiles (4 float4 loads per 16 FMAs):
//   for j = 0..3:  (a) block column j -= L[:, :32j] L[32j.., :32j]^T                (left-looking, K = 32 j)
//                  (b) warp 0: L_jj = chol(diagonal block), D_j = inv(L_jj)          (lane = row, then lane = column)
//                  (c) rows below: L[:, j] = X[:, j] D_j^T                             (thread = row)
//   for r = 1..3:  S = L[r, :r] W[:r, :r];  W[r, :r] = -D_r S                         (block row r of W = inv(L11))
// W[i][j] (i >= j) lives TRANSPOSED in the unused upper triangle of the same tile, at tile[j][i + 4]: both factors
// of every product are then contiguous in k.  256 threads, 100 KB of shared memory, two CTAs per SM.
// INV_ONLY (blocked LU): the block is a finished unit-lower L11; warps 0..3 invert the four diagonal blocks, then
// the block rows of W follow as above.  Plain fp32 FMA arithmetic (tensor paths: tolerance-checked).
// ------------------------------------------------------------------------------------------------
struct Diag2 {
    static constexpr int LDT = 132;  // row stride of the tile: rows stay 16-byte aligned, W^T fits with its shift
    static constexpr int SH = 4;     // W[i][j] is kept at tile[j][i + SH]
    static constexpr int LDN = 36;   // row stride of the inverted diagonal blocks (natural layout) and of the scratch
    static constexpr int kTile = NB * LDT, kDn = 4 * 32 * LDN, kSt = 96 * LDN;
    static constexpr int kBytes = (kTile + kDn + kSt) * 4;
};

__device__ __forceinline__ float dot4(float4 a, float4 b, float s) {
    s = fmaf(a.x, b.x, s);
    s = fmaf(a.y, b.y, s);
    s = fmaf(a.z, b.z, s);
    return fmaf(a.w, b.w, s);
}

// One warp: Cholesky of the 32 x 32 diagonal block at (c0, c0) of the tile.  lane = row; the row lives in registers,
// the pivot and the column stream travel by shuffle.  L_jj goes back into the tile (lower part only: W^T lives above
// it), 1 / l_kk into rd[0..31].  This is the serial part of the kernel (four times 32 dependent pivot steps), so it
// holds nothing else in registers: the inverse of the block is taken later, by other warps (inv32).
__device__ __forceinline__ void chol32(float* T, float* rd, int c0, int lane) {
    constexpr int LDT = Diag2::LDT;
    float d[32];
    {
        const float4* row4 = reinterpret_cast<const float4*>(T + (c0 + lane) * LDT + c0);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 v = row4[q];
            d[4 * q] = v.x, d[4 * q + 1] = v.y, d[4 * q + 2] = v.z, d[4 * q + 3] = v.w;
        }
    }
    static_for<0, 32>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const float pk = __shfl_sync(0xffffffffu, d[k], k);
        const float rk = rsqrtf(pk);  // 2 ulp: far inside the tolerance of the tensor path
        if (lane == k) rd[k] = rk;
        const float lk = d[k] * rk;   // l_ik for rows i > k, sqrt(a_kk) in lane k
        d[k] = lk;
#pragma unroll
        for (int c = k + 1; c < 32; c++) d[c] = fmaf(-lk, __shfl_sync(0xffffffffu, lk, c), d[c]);
    });
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (c <= lane) T[(c0 + lane) * LDT + c0 + c] = d[c];
}

// One warp: D = inv(L_jj) of the 32 x 32 lower-triangular block at (c0, c0) of the tile; lane = column of D:
// x_i = (delta_i,lane - sum_{k<i} l_ik x_k) * rd_i, rows of L_jj read as broadcast float4 loads.  D goes to Dn (natural
// layout, zeros above the diagonal) and, transposed, into the upper triangle of the tile (W_jj).
__device__ __forceinline__ void inv32(float* T, float* Dn, const float* rd, int c0, int lane) {
    constexpr int LDT = Diag2::LDT, LDN = Diag2::LDN, SH = Diag2::SH;
    float x[32];
    static_for<0, 32>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        float a0 = (i == lane) ? 1.0f : 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        const float4* lrow = reinterpret_cast<const float4*>(T + (c0 + i) * LDT + c0);
        static_for<0, (i + 3) / 4>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const float4 lv = lrow[q];
            if constexpr (4 * q < i) a0 = fmaf(-lv.x, x[4 * q], a0);
            if constexpr (4 * q + 1 < i) a1 = fmaf(-lv.y, x[4 * q + 1], a1);
            if constexpr (4 * q + 2 < i) a2 = fmaf(-lv.z, x[4 * q + 2], a2);
            if constexpr (4 * q + 3 < i) a3 = fmaf(-lv.w, x[4 * q + 3], a3);
        });
        x[i] = ((a0 + a1) + (a2 + a3)) * rd[i];
    });
#pragma unroll
    for (int i = 0; i < 32; i++) {
        Dn[i * LDN + lane] = x[i];                                  // (x_i = 0 for i < lane)
        if (i >= lane) T[(c0 + lane) * LDT + c0 + i + SH] = x[i];  // W_jj, transposed into the upper triangle
    }
}

template <bool INV_ONLY>
__global__ void __launch_bounds__(256, 2)
blk_diag2_kernel(const float* __restrict__ S, float* __restrict__ L, float* __restrict__ W, int n, int o,
                 const int* __restrict__ rowmap = nullptr) {
    constexpr int LDT = Diag2::LDT, LDN = Diag2::LDN, SH = Diag2::SH;
    extern __shared__ __align__(16) float d2_smem[];
    float* T = d2_smem;
    float* Dn = T + Diag2::kTile;  // [4][32][LDN]
    float* St = Dn + Diag2::kDn;   // [96][LDN]: S transposed (column-major), S[i][c] at St[c][i]
    __shared__ float rdiag[NB];    // 1 / l_kk (1 for the unit-lower block of the LU)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* sblk = S + (long)blockIdx.x * n * n + (long)o * n + o;  // S = A for block column 0, L afterwards
    float* blk = L + (long)blockIdx.x * n * n + (long)o * n + o;
    // rowmap (blocked LU, lu_crout.cu): row r of the block is row rowmap[o + r] of the matrix S (rows are never moved there)
    const int* rmap = rowmap ? rowmap + (long)blockIdx.x * n + o : nullptr;
    for (int e = tid; e < NB * (NB / 4); e += 256) {
        const int row = e >> 5, c4 = e & 31;
        const float* srow = rmap ? S + (long)blockIdx.x * n * n + (long)rmap[row] * n + o : sblk + (long)row * n;
        *reinterpret_cast<float4*>(T + row * LDT + 4 * c4) = *(reinterpret_cast<const float4*>(srow) + c4);
    }
    if (INV_ONLY && tid < NB) rdiag[tid] = 1.0f;
    __syncthreads();
    if constexpr (!INV_ONLY) {
        for (int j = 0; j < 4; j++) {
            const int c0 = 32 * j;
            if (j > 0) {
                // (a) left-looking update of block column j: 2 x 2 tiles (rows r0, r0+1; columns ca, ca+16), K = c0
                const int tiles = ((NB - c0) / 2) * 16;
                for (int t = tid; t < tiles; t += 256) {
                    const int cp = t & 15, rp = t >> 4;
                    const int r0 = c0 + 2 * rp, ca = c0 + cp, cb = ca + 16;
                    if (r0 + 1 < ca) continue;  // entirely above the diagonal
                    const float4* A0 = reinterpret_cast<const float4*>(T + r0 * LDT);
                    const float4* A1 = reinterpret_cast<const float4*>(T + (r0 + 1) * LDT);
                    const float4* B0 = reinterpret_cast<const float4*>(T + ca * LDT);
                    const float4* B1 = reinterpret_cast<const float4*>(T + cb * LDT);
                    float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
                    for (int k4 = 0; k4 < c0 / 4; k4++) {
                        const float4 a0 = A0[k4], a1 = A1[k4], b0 = B0[k4], b1 = B1[k4];
                        s00 = dot4(a0, b0, s00), s01 = dot4(a0, b1, s01);
                        s10 = dot4(a1, b0, s10), s11 = dot4(a1, b1, s11);
                    }
                    if (r0 >= ca) T[r0 * LDT + ca] -= s00;
                    if (r0 >= cb) T[r0 * LDT + cb] -= s01;
                    if (r0 + 1 >= ca) T[(r0 + 1) * LDT + ca] -= s10;
                    if (r0 + 1 >= cb) T[(r0 + 1) * LDT + cb] -= s11;
                }
                __syncthreads();
            }
            // (b) the 32 x 32 diagonal block: Cholesky inside warp 0 (the serial part).  Meanwhile the other seven warps
            //     write a quarter of the zeros this block row of L needs right of its diagonal block (nothing else ever
            //     writes those blocks; the stores are fire-and-forget and cost the factorisation nothing)
            if (warp == 0) {
                chol32(T, rdiag + c0, c0, lane);
            } else {
                const int zcols4 = (n - o - NB) / 4;  // float4 per row right of the diagonal block
                float4* zbase = reinterpret_cast<float4*>(blk + NB);
                for (int row = c0; row < c0 + 32; row++)
                    for (int c4 = tid - 32; c4 < zcols4; c4 += 224)
                        __stcs(zbase + (long)row * (n / 4) + c4, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            }
            __syncthreads();
            // (c) rows below the diagonal block, thread = row, forward substitution against L_jj (read as broadcast
            //     float4 loads): l_rc = (x_c - sum_{m<c} l_rm l_cm) / l_cc
            if (tid < NB - c0 - 32) {
                float* xrow = T + (c0 + 32 + tid) * LDT + c0;
                float xr[32];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 v = reinterpret_cast<const float4*>(xrow)[q];
                    xr[4 * q] = v.x, xr[4 * q + 1] = v.y, xr[4 * q + 2] = v.z, xr[4 * q + 3] = v.w;
                }
                static_for<0, 32>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    float a0 = xr[c], a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                    const float4* lrow = reinterpret_cast<const float4*>(T + (c0 + c) * LDT + c0);
                    static_for<0, (c + 3) / 4>([&](auto qc) {
                        constexpr int q = decltype(qc)::value;
                        const float4 lv = lrow[q];
                        if constexpr (4 * q < c) a0 = fmaf(-lv.x, xr[4 * q], a0);
                        if constexpr (4 * q + 1 < c) a1 = fmaf(-lv.y, xr[4 * q + 1], a1);
                        if constexpr (4 * q + 2 < c) a2 = fmaf(-lv.z, xr[4 * q + 2], a2);
                        if constexpr (4 * q + 3 < c) a3 = fmaf(-lv.w, xr[4 * q + 3], a3);
                    });
                    xr[c] = ((a0 + a1) + (a2 + a3)) * rdiag[c0 + c];
                });
#pragma unroll
                for (int q = 0; q < 8; q++)
                    reinterpret_cast<float4*>(xrow)[q] = make_float4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
            }
            __syncthreads();
        }
    }
    // D_j = inv(L_jj) for the four diagonal blocks, one warp each (off the serial path: nothing above needed them)
    if (warp >= 4) inv32(T, Dn + (warp - 4) * 32 * LDN, rdiag + 32 * (warp - 4), 32 * (warp - 4), lane);
    __syncthreads();
    // block rows 1..3 of W = inv(L11)
    for (int r = 1; r < 4; r++) {
        const int r0 = 32 * r;
        // S[i][c] = sum_{k = c}^{r0 - 1} L[r0 + i][k] W[k][c]: tiles (rows ia, ia + 16) x (columns ca, ca + 1)
        for (int t = tid; t < 16 * (r0 / 2); t += 256) {
            const int ip = t & 15, cp = t >> 4;
            const int ca = 2 * cp, ia = ip, ib = ip + 16;
            const float4* A0 = reinterpret_cast<const float4*>(T + (r0 + ia) * LDT);
            const float4* A1 = reinterpret_cast<const float4*>(T + (r0 + ib) * LDT);
            const float4* B0 = reinterpret_cast<const float4*>(T + ca * LDT + SH);        // W[k][ca] at k
            const float4* B1 = reinterpret_cast<const float4*>(T + (ca + 1) * LDT + SH);  // W[k][ca + 1]
            float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
            const int kfirst = ca >> 2;
            for (int k4 = kfirst; k4 < r0 / 4; k4++) {
                const float4 a0 = A0[k4], a1 = A1[k4];
                float4 b0 = B0[k4], b1 = B1[k4];
                if (k4 == kfirst) {  // W[k][c] = 0 for k < c: what sits there in the tile is L
                    const int k = 4 * k4;
                    if (k < ca) b0.x = 0.0f;
                    if (k + 1 < ca) b0.y = 0.0f;
                    if (k + 2 < ca) b0.z = 0.0f;
                    if (k < ca + 1) b1.x = 0.0f;
                    if (k + 1 < ca + 1) b1.y = 0.0f;
                    if (k + 2 < ca + 1) b1.z = 0.0f;
                    if (k + 3 < ca + 1) b1.w = 0.0f;
                }
                s00 = dot4(a0, b0, s00), s01 = dot4(a0, b1, s01);
                s10 = dot4(a1, b0, s10), s11 = dot4(a1, b1, s11);
            }
            St[ca * LDN + ia] = s00, St[(ca + 1) * LDN + ia] = s01;
            St[ca * LDN + ib] = s10, St[(ca + 1) * LDN + ib] = s11;
        }
        __syncthreads();
        // W[r0 + i][c] = -sum_{m <= i} D_r[i][m] S[m][c]
        const float* dn = Dn + r * 32 * LDN;
        for (int t = tid; t < 16 * (r0 / 2); t += 256) {
            const int ip = t & 15, cp = t >> 4;
            const int ca = 2 * cp, ia = ip, ib = ip + 16;
            const float4* D0 = reinterpret_cast<const float4*>(dn + ia * LDN);
            const float4* D1 = reinterpret_cast<const float4*>(dn + ib * LDN);
            const float4* S0 = reinterpret_cast<const float4*>(St + ca * LDN);
            const float4* S1 = reinterpret_cast<const float4*>(St + (ca + 1) * LDN);
            float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
            for (int m4 = 0; m4 <= (ib >> 2); m4++) {  // D_r[i][m] = 0 for m > i (stored)
                const float4 d0 = D0[m4], d1 = D1[m4], v0 = S0[m4], v1 = S1[m4];
                s00 = dot4(d0, v0, s00), s01 = dot4(d0, v1, s01);
                s10 = dot4(d1, v0, s10), s11 = dot4(d1, v1, s11);
            }
            T[ca * LDT + r0 + ia + SH] = -s00, T[(ca + 1) * LDT + r0 + ia + SH] = -s01;
            T[ca * LDT + r0 + ib + SH] = -s10, T[(ca + 1) * LDT + r0 + ib + SH] = -s11;
        }
        __syncthreads();
    }
    if constexpr (!INV_ONLY) {
        for (int e = tid; e < NB * (NB / 4); e += 256) {
            const int row = e >> 5, c4 = e & 31;
            float4 v = *reinterpret_cast<const float4*>(T + row * LDT + 4 * c4);
            if (4 * c4 > row) v.x = 0.0f;
            if (4 * c4 + 1 > row) v.y = 0.0f;
            if (4 * c4 + 2 > row) v.z = 0.0f;
            if (4 * c4 + 3 > row) v.w = 0.0f;
            *(reinterpret_cast<float4*>(blk + (long)row * n) + c4) = v;
        }
    }
    float* wout = W + (long)blockIdx.x * NB * NB;
    for (int e = tid; e < NB * NB; e += 256) {
        const int i = e >> 7, j = e & (NB - 1);
        wout[e] = (j <= i) ? T[j * LDT + i + SH] : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// tensor-core tile kernel
// ------------------------------------------------------------------------------------------------

// TRSM tiles of the blocked Cholesky: L[r][q] = X[r][q] * W^T.  grid: (row blocks below the diagonal block, batch)
__global__ void __launch_bounds__(256, MmaSmem::kStages == 1 ? 3 : 2)
blk_mma_kernel(const float* __restrict__ S, float* __restrict__ L, const float* __restrict__ W, int n, int o) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space: LDS/STS, not generic LD/ST)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MmaSmem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MmaSmem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float* Lb = L + (long)blockIdx.y * n * n;
    const float* Sb = S + (long)blockIdx.y * n * n;  // S = A for panel 0 (inputs not yet in L), L afterwards
    const int tj = blockIdx.x;
    const int row0 = o + NB;  // first row of the panel below the diagonal block
    // operand rows (K-major, k = panel columns o .. o+127): the not yet scaled panel rows X, and W (B[i][k] = W[i][k])
    const float* Arows = Sb + (long)(row0 + tj * NB) * n + o;
    const long lda = n;
    const float* Brows = W + (long)blockIdx.y * NB * NB;
    const long ldb = NB;

    if (warp == 0) tmem_alloc(tmem_ptr, 128);
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    constexpr int kChunks = NB / MmaSmem::kChunkK;  // 4 chunks of 32 k
    // 16-byte piece `it` of this thread: row = idx / 8, chunk-of-16-bytes c4 = idx % 8 (SWIZZLE_128B: c4 ^= row % 8)
    auto piece_off = [&](int it) {
        const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
        return row * 128 + ((c4 ^ (row & 7)) << 4);
    };
    auto issue_loads = [&](int chunk) {
        unsigned char* ra = smem + MmaSmem::kRawA + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
        unsigned char* rb = smem + MmaSmem::kRawB + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
            cp_async16(ra + piece_off(it), Arows + row * lda + chunk * MmaSmem::kChunkK + c4 * 4);
            cp_async16(rb + piece_off(it), Brows + row * ldb + chunk * MmaSmem::kChunkK + c4 * 4);
        }
        cp_async_commit();
    };
    issue_loads(0);
    for (int chunk = 0; chunk < kChunks; chunk++) {
        unsigned char* ra = smem + MmaSmem::kRawA + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
        unsigned char* rb = smem + MmaSmem::kRawB + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
        unsigned char* la = smem + MmaSmem::kLoA;
        unsigned char* lb = smem + MmaSmem::kLoB;
        // the other raw stage was last read by the MMAs of chunk-1, which have completed (mbar_wait below)
        if (MmaSmem::kStages == 2 && chunk + 1 < kChunks) {
            issue_loads(chunk + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        // each thread splits exactly the pieces it copied itself, so no barrier is needed before reading them
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int off = piece_off(it);
            *reinterpret_cast<float4*>(la + off) = lo_of(*reinterpret_cast<const float4*>(ra + off));
            *reinterpret_cast<float4*>(lb + off) = lo_of(*reinterpret_cast<const float4*>(rb + off));
        }
        fence_async_smem();  // cp.async / st.shared data -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (t == 0) {
            const uint64_t dAhi = make_sw128_desc(smem_u32(ra)), dAlo = make_sw128_desc(smem_u32(la));
            const uint64_t dBhi = make_sw128_desc(smem_u32(rb)), dBlo = make_sw128_desc(smem_u32(lb));
#pragma unroll
            for (int kk = 0; kk < MmaSmem::kChunkK / 8; kk++) {
                // k-step kk: 8 tf32 = 32 bytes further inside the 128-byte swizzled row
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t first = (chunk == 0 && kk == 0) ? 0u : 1u;
                umma_tf32(tmem_d, dAlo + adv, dBhi + adv, kIdescTf32, first);
                umma_tf32(tmem_d, dAhi + adv, dBlo + adv, kIdescTf32, 1u);
                umma_tf32(tmem_d, dAhi + adv, dBhi + adv, kIdescTf32, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, (uint32_t)(chunk & 1));  // MMAs done: lo tiles and this raw stage are free again
        tc_fence_after();
        if (MmaSmem::kStages == 1 && chunk + 1 < kChunks) issue_loads(chunk + 1);
    }

    // epilogue.  TMEM -> registers (thread = accumulator row 32*(warp%4)+lane, 64 columns (warp/4)) ->
    // shared memory (the operand tiles are free now; 128 x 128 fp32 = 64 KB, 16-byte pieces XOR-swizzled
    // by row so that both the row-per-lane writes and the row-per-warp reads are conflict-free) ->
    // coalesced stores, one 512-byte row per warp instruction.
    {
        float* stage = reinterpret_cast<float*>(smem);
        const int drow = (warp & 3) * 32 + lane;
        const int dcol0 = (warp >> 2) * 64;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float acc[32];
            const int col = dcol0 + h * 32;
            tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, acc);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int c4 = (col >> 2) + q;  // 16-byte piece index inside the row (0..31)
                *reinterpret_cast<float4*>(stage + drow * 128 + ((c4 ^ (drow & 7)) << 2)) =
                    make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        __syncthreads();
        float* Cbase = Lb + (long)(row0 + tj * NB) * n + o;  // L21
#pragma unroll 4
        for (int rr = 0; rr < 16; rr++) {
            const int row = warp * 16 + rr;
            const float4 d = *reinterpret_cast<const float4*>(stage + row * 128 + ((lane ^ (row & 7)) << 2));
            *(reinterpret_cast<float4*>(Cbase + (long)row * n) + lane) = d;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 128);
}

// ------------------------------------------------------------------------------------------------
// Left-looking trailing update: tile (r, q) of block column q receives ALL earlier panels at once,
//      X[r][q] = A[r][q] - sum_{j<q} L[r][j] * L[q][j]^T         (K = 128 q, up to 896)
// so C is read (from A) and written (to L) once per tile instead of once per panel, and the K loop is
// long enough to pipeline: global loads run two chunks ahead in registers, the threads write the raw
// (= hi) and lo tiles of chunk c while the tensor core still multiplies chunk c-1 (both double
// buffered), completion of chunk c-2 (tcgen05.commit -> one of two mbarriers) frees its buffers.
// (A cp.async ring would cost a third pass over shared memory - write raw, read raw, write lo.)
// Two row blocks per CTA: tiles (r0, q) and (r0 + 1, q) share the B operand (block row q), which cuts
// the panel bytes each step re-reads from L2/HBM by a quarter - the measured bound of the update:
// three operand tiles per chunk (A0, A1, B: raw + lo, double buffered = 192 KB).  The shared tile is the M operand and
// the two row tiles - contiguous in shared memory - one 256-row N operand, so ONE 128 x 256 accumulator holds the two
// results TRANSPOSED (lane = column of block column q): the epilogue reads C from A and writes X to L straight from
// registers, coalesced over the columns, without a shared-memory stage.  grid: (ceil((panels - q) / 2), batch)
// ------------------------------------------------------------------------------------------------
struct Ll2Smem {
    static constexpr int kTile = 128 * 32 * 4;
    static constexpr int kRaw = 0;           // [2 stages][A0, A1, B]
    static constexpr int kLo = 6 * kTile;    // [2 stages][A0, A1, B]
    static constexpr int kBar = 12 * kTile;
    static constexpr int kTmemPtr = kBar + 16;
    static constexpr int kTotal = kBar + 32 + 1024;
};

__global__ void __launch_bounds__(512, 1)
blk_ll2_kernel(const float* __restrict__ S, float* __restrict__ L, int n, int q) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Ll2Smem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Ll2Smem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int panels = n / NB;
    const int r0 = q + 2 * blockIdx.x;
    const bool two = r0 + 1 < panels;  // the last CTA of an odd column holds one tile only
    float* Lb = L + (long)blockIdx.y * n * n;
    const float* Sb = S + (long)blockIdx.y * n * n;
    const float* A0rows = Lb + (long)r0 * NB * n;
    const float* A1rows = A0rows + (long)NB * n;
    const float* Brows = Lb + (long)q * NB * n;
    const bool same = (r0 == q);  // tile 0 is the diagonal tile: B is A0
    const int C = 4 * q;

    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    auto piece_off = [&](int it) {
        const int idx = it * 512 + t, row = idx >> 3, c4 = idx & 7;
        return row * 128 + ((c4 ^ (row & 7)) << 4);
    };
    float4 b0[2][2], b1[2][2], bb[2][2];  // [register stage][piece]: A0, A1, B
    auto gload = [&](int c, int sreg) {
        if (c < C) {
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int idx = it * 512 + t, row = idx >> 3, c4 = idx & 7;
                const long off = (long)row * n + c * 32 + c4 * 4;
                b0[sreg][it] = __ldg(reinterpret_cast<const float4*>(A0rows + off));
                if (two) b1[sreg][it] = __ldg(reinterpret_cast<const float4*>(A1rows + off));
                if (!same) bb[sreg][it] = __ldg(reinterpret_cast<const float4*>(Brows + off));
            }
        }
    };
    auto step = [&](int c, int sreg) {
        if (c >= 2) {
            mbar_wait(&bars[c & 1], (uint32_t)(((c - 2) >> 1) & 1));
            tc_fence_after();
        }
        unsigned char* ra0 = smem + Ll2Smem::kRaw + (c & 1) * 3 * Ll2Smem::kTile;
        unsigned char* la0 = smem + Ll2Smem::kLo + (c & 1) * 3 * Ll2Smem::kTile;
        unsigned char* ra1 = ra0 + Ll2Smem::kTile;
        unsigned char* la1 = la0 + Ll2Smem::kTile;
        unsigned char* rb = same ? ra0 : ra0 + 2 * Ll2Smem::kTile;
        unsigned char* lb = same ? la0 : la0 + 2 * Ll2Smem::kTile;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int off = piece_off(it);
            *reinterpret_cast<float4*>(ra0 + off) = b0[sreg][it];
            *reinterpret_cast<float4*>(la0 + off) = lo_of(b0[sreg][it]);
            if (two) {
                *reinterpret_cast<float4*>(ra1 + off) = b1[sreg][it];
                *reinterpret_cast<float4*>(la1 + off) = lo_of(b1[sreg][it]);
            }
            if (!same) {
                *reinterpret_cast<float4*>(rb + off) = bb[sreg][it];
                *reinterpret_cast<float4*>(lb + off) = lo_of(bb[sreg][it]);
            }
        }
        gload(c + 2, sreg);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (t == 0) {
            // ONE accumulator of 128 x 256 (transposed result): block row q's tile is the M operand, the two paired row
            // tiles - contiguous in shared memory - form one 256-row N operand, so the shared tile is read once per
            // k-step instead of twice (shared-memory bandwidth bounds this kernel: 12 KB instead of 16 KB of operand
            // reads per pair of 128 x 128 x 8 products)
            const uint64_t dMhi = make_sw128_desc(smem_u32(rb)), dMlo = make_sw128_desc(smem_u32(lb));
            const uint64_t dNhi = make_sw128_desc(smem_u32(ra0)), dNlo = make_sw128_desc(smem_u32(la0));
            const uint32_t idesc = two ? kIdescTf32N256 : kIdescTf32;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t first = (c == 0 && kk == 0) ? 0u : 1u;
                umma_tf32(tmem_d, dMlo + adv, dNhi + adv, idesc, first);
                umma_tf32(tmem_d, dMhi + adv, dNlo + adv, idesc, 1u);
                umma_tf32(tmem_d, dMhi + adv, dNhi + adv, idesc, 1u);
            }
            umma_commit(&bars[c & 1]);
        }
    };
    gload(0, 0);
    gload(1, 1);
    for (int c = 0; c < C; c += 2) {  // C = 4 q is even
        step(c, 0);
        step(c + 1, 1);
    }
    // epilogue.  The accumulator holds the TRANSPOSED tiles: lane = column c of block column q, accumulator column
    // h * 128 + j = row j of row tile r0 + h.  For a fixed row the 32 lanes of a warp cover 32 consecutive columns, so C
    // is read from A and X written to L with coalesced 128-byte accesses straight from / to registers.  The C values are
    // fetched while the last chunks are still being multiplied.
    const int dlane = (warp & 3) * 32 + lane;  // column within the block column
    const int rbase = (warp >> 2) * 32;        // 32 rows of each tile
    const long coff = (long)r0 * NB * n + (long)q * NB + dlane;
    float cv[2][32];
#pragma unroll
    for (int h = 0; h < 2; h++)
        if (h == 0 || two) {
#pragma unroll
            for (int j = 0; j < 32; j++) cv[h][j] = __ldg(Sb + coff + (long)(h * NB + rbase + j) * n);
        }
    mbar_wait(&bars[(C - 2) & 1], (uint32_t)(((C - 2) >> 1) & 1));
    mbar_wait(&bars[(C - 1) & 1], (uint32_t)(((C - 1) >> 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (h == 1 && !two) break;
        float acc[32];
        tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * 128 + rbase), acc);
#pragma unroll
        for (int j = 0; j < 32; j++) Lb[coff + (long)(h * NB + rbase + j) * n] = cv[h][j] - acc[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// left-looking driver: per block column q: update (if q > 0), diagonal block, TRSM
static cudaError_t cholesky_left_looking(const float* A, float* L, float* W, int n, long batch, cudaStream_t st) {
    cudaError_t e;
    if ((e = ensure_dynamic_smem(blk_mma_kernel, MmaSmem::kTotal)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(blk_ll2_kernel, Ll2Smem::kTotal)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(blk_diag_kernel<DIAG_CHOL_INV>, kDiagSmem)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(blk_diag2_kernel<false>, Diag2::kBytes)) != cudaSuccess) return e;
    const bool old_diag = option(HLSD_OPT_DIAG_KERNEL) == 1;  // the round-1 column-by-column kernel (A/B measurements)
    const int panels = n / NB;
    const unsigned nb = (unsigned)batch;
    if (old_diag) {  // (blk_diag2_kernel writes the zero blocks of its block row itself)
        ProfScope ps(HLSD_PROF_OTHER, st);
        blk_zero_upper_kernel<<<dim3(panels * (panels - 1) / 2, nb), 256, 0, st>>>(L, n);
        count_launch();
    }
    for (int q = 0; q < panels; q++) {
        const int o = q * NB;
        const int mt = panels - q - 1;
        const float* src = (q == 0) ? A : L;  // where block column q currently lives
        if (q > 0) {
            {
                ProfScope ps(HLSD_PROF_UPDATE, st);
                blk_ll2_kernel<<<dim3((panels - q + 1) / 2, nb), 512, Ll2Smem::kTotal, st>>>(A, L, n, q);
                count_launch();
            }
        }
        {
            ProfScope ps(HLSD_PROF_DIAG, st);
            if (old_diag) blk_diag_kernel<DIAG_CHOL_INV><<<nb, 512, kDiagSmem, st>>>(src, L, W, n, o);
            else blk_diag2_kernel<false><<<nb, 256, Diag2::kBytes, st>>>(src, L, W, n, o);
            count_launch();
        }
        if (mt > 0) {
            {
                ProfScope ps(HLSD_PROF_TRSM, st);
                blk_mma_kernel<<<dim3(mt, nb), 256, MmaSmem::kTotal, st>>>(src, L, W, n, o);
                count_launch();
            }
        }
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Bit-identical Cholesky of 128 x 128 matrices (right-looking designs): the register-tiled diagonal-block
// kernel in exact arithmetic, one CTA per matrix (fused != 0: with the contraction HLSD_FLAG_FUSED allows).
cudaError_t cholesky128_exact(const float* A, float* L, long batch, int fused, cudaStream_t st) {
    auto kern = fused ? blk_diag_kernel<DIAG_EXACT, true> : blk_diag_kernel<DIAG_EXACT, false>;
    cudaError_t e = ensure_dynamic_smem(kern, kDiagSmem);
    if (e != cudaSuccess) return e;
    for (long b0 = 0; b0 < batch; b0 += 1 << 20) {
        const long cnt = std::min<long>(batch - b0, 1 << 20);
        kern<<<(unsigned)cnt, 512, kDiagSmem, st>>>(A + b0 * NB * NB, L + b0 * NB * NB, nullptr, NB, 0);
        count_launch();
    }
    return cudaGetLastError();
}

// entry points for the exact blocked path (blocked_exact.cu): the bit-identical 128 x 128 kernel on the diagonal
// block at (o, o) of every matrix, and the zero fill of the blocks above the block diagonal
cudaError_t chol_diag_exact(const float* S, float* L, int n, int o, long batch, cudaStream_t st) {
    auto kern = blk_diag_kernel<DIAG_EXACT, false>;
    cudaError_t e = ensure_dynamic_smem(kern, kDiagSmem);
    if (e != cudaSuccess) return e;
    ProfScope ps(HLSD_PROF_DIAG, st);
    kern<<<(unsigned)batch, 512, kDiagSmem, st>>>(S, L, nullptr, n, o);
    count_launch();
    return cudaGetLastError();
}
cudaError_t chol_zero_upper_blocks(float* L, int n, long batch, cudaStream_t st) {
    const int panels = n / NB;
    ProfScope ps(HLSD_PROF_OTHER, st);
    blk_zero_upper_kernel<<<dim3(panels * (panels - 1) / 2, (unsigned)batch), 256, 0, st>>>(L, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t cholesky_blocked_exact(const float* A, float* L, int n, long batch, cudaStream_t st);  // blocked_exact.cu
cudaError_t cholesky_blocked_exact_v22(const float* A, float* L, int n, long batch, cudaStream_t st);

cudaError_t cholesky_blocked(const float* A, float* L, int n, long batch, int use_tensor, int variant, cudaStream_t st) {
    if (variant == 1) return use_tensor ? cudaErrorNotSupported : cholesky_blocked_exact_v22(A, L, n, batch, st);
    if (!use_tensor) return cholesky_blocked_exact(A, L, n, batch, st);
    if (n % NB != 0 || n < 2 * NB) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    if (batch > 65535) return cudaErrorNotSupported;
    float* W = nullptr;  // inverted diagonal blocks of the current block column, one per matrix
    cudaError_t e = ws_alloc(reinterpret_cast<void**>(&W), sizeof(float) * (size_t)batch * NB * NB, st);
    if (e != cudaSuccess) return e;
    e = cholesky_left_looking(A, L, W, n, batch, st);
    ws_free(W, st);
    return e;
}

// =================================================================================================
// Blocked LU with partial pivoting in the reference's EXACT arithmetic (HLSD_PATH_BLOCKED), 256 <= N <= 1024, N a
// multiple of 128 - the kernels below and blocked_exact.cu.  (The tensor-core LU lives in lu_crout.cu; it shares
// blk_diag2_kernel<INV_ONLY> through lu_inverse_l11.)
//
// Same elimination as the reference's LU designs (LU/lu_v1.1/model4x4/lu.cpp:23-92: pivot = first row with the largest
// |a_ik|, rows exchanged in columns >= k only, multipliers l_ik = a_ik / a_kk, a_ij -= l_ik u_kj), right-looking over
// 128-column panels, every update applied in ascending k so that each element sees the reference's operations in the
// reference's order:
//   per panel p (columns o .. o+127), all matrices of the batch at once:
//     1. lu_panel_kernel<true> (SIMT) pivoted elimination of the (N-o) x 128 panel, thread = row, in four 32-column
//                          sub-panels held in registers; writes L (unit lower) to the L buffer, U11 to the working
//                          matrix, the pivots to P
//     2. lu_swap_kernel    (SIMT) the panel's row exchanges applied to the columns right of the panel
//     3. xlu_u12_and_update (blocked_exact.cu) U12 by forward substitution, then the trailing update on SIMT tiles
//     4. lu_unswap_panel_kernel (SIMT) inside a panel rows are exchanged whole (the deferred block updates need L in
//                          final row order); once L21 has served as operand, the panel's L columns go back to the
//                          reference's convention (multipliers stay where they were computed)
// The working matrix is the U output buffer (a copy of A); what remains in it at the end is U.
// =================================================================================================
struct PivCand {
    float v;
    int i;
};
__device__ __forceinline__ PivCand piv_better(PivCand a, PivCand b) {
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

// Panel factorisation: one CTA per matrix, thread t = row o + t of the working matrix (blockDim = rows of
// the panel, at least 256 because threads 0..191 also serve as column workers between sub-panels).
// EXACT: the reference's arithmetic (IEEE division, separately rounded multiply and subtract); the per-element
// operation order is the reference's in both modes (every update arrives in ascending k), so with EXACT the panel -
// and with the exact U12 / trailing-update kernels of blocked_exact.cu the whole factorisation - is bit-identical.
template <bool EXACT>
__global__ void __launch_bounds__(1024, 1)
lu_panel_kernel(float* __restrict__ Wk, float* __restrict__ L, int* __restrict__ P, int n, int o, int pivoting) {
    // double buffered by column parity: two barriers per column are enough (after the search partials, after the
    // pivot row is published); a warp that runs ahead into the next column writes the other buffers
    __shared__ float urow_b[2][32];  // pivot row, columns kk.. of the sub-panel
    __shared__ float krow_b[2][32];  // the row that sat at position k before the exchange
    __shared__ PivCand red_b[2][32];
    __shared__ int spiv[32];         // pivots of the current sub-panel (absolute row indices)
    __shared__ float l11[32][33];    // multipliers of the sub-panel's own 32 rows
    __shared__ __align__(16) float ublk[32][96];  // U rows of the sub-panel for the columns right of it
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int R = n - o;             // rows of the panel
    const int row = o + t;
    const bool live = t < R;
    float* wk = Wk + (long)blockIdx.x * n * n;
    float* lm = L + (long)blockIdx.x * n * n;
    int* pv = P + (long)blockIdx.x * n;
    for (int sp = 0; sp < 4; sp++) {
        const int cs = o + 32 * sp;  // first column of the sub-panel
        float a[32];
        if (live) {
            const float4* src = reinterpret_cast<const float4*>(wk + (long)row * n + cs);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float4 v = src[q];
                a[4 * q] = v.x, a[4 * q + 1] = v.y, a[4 * q + 2] = v.z, a[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 32; q++) a[q] = 0.0f;
        }
        static_for<0, 32>([&](auto kkc) {
            constexpr int kk = decltype(kkc)::value;
            const int k = cs + kk;  // pivot position (absolute row = absolute column)
            int p = k;
            float* urow = urow_b[kk & 1];
            float* krow = krow_b[kk & 1];
            if (k < n - 1 && pivoting) {  // (HLSD_LU_NOPIVOT: P[k] = k, no search)
                PivCand* red = red_b[kk & 1];
                PivCand c{-2.0f, 0x7fffffff};
                if (live && row >= k) {
                    float v = fabsf(a[kk]);
                    if (v != v) v = (row == k) ? __int_as_float(0x7f800000) : -1.0f;
                    c = PivCand{v, row};
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    PivCand o2{__shfl_xor_sync(0xffffffffu, c.v, off), __shfl_xor_sync(0xffffffffu, c.i, off)};
                    c = piv_better(c, o2);
                }
                if (lane == 0) red[warp] = c;
                __syncthreads();
                // every warp reduces the per-warp candidates itself (no second barrier, no broadcast)
                PivCand c2 = lane < (int)(blockDim.x >> 5) ? red[lane] : PivCand{-2.0f, 0x7fffffff};
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    PivCand o2{__shfl_xor_sync(0xffffffffu, c2.v, off), __shfl_xor_sync(0xffffffffu, c2.i, off)};
                    c2 = piv_better(c2, o2);
                }
                p = c2.i;
            }
            if (t == 0) {
                spiv[kk] = p;
                pv[k] = p;
            }
            // whole rows are exchanged (multipliers included, LAPACK style: the deferred block updates below
            // need L in final row order; lu_unswap_panel_kernel restores the reference's convention afterwards)
            if (live && row == p) {
#pragma unroll
                for (int j = 0; j < 32; j++) urow[j] = a[j];
            }
            if (live && row == k && p != k) {
#pragma unroll
                for (int j = 0; j < 32; j++) krow[j] = a[j];
            }
            __syncthreads();
            if (live && p != k) {
                if (row == k) {
#pragma unroll
                    for (int j = 0; j < 32; j++) a[j] = urow[j];
                } else if (row == p) {
#pragma unroll
                    for (int j = 0; j < 32; j++) a[j] = krow[j];
                }
            }
            if (live && row > k) {
                const float l = EXACT ? xdiv(a[kk], urow[kk]) : a[kk] / urow[kk];
                a[kk] = l;
#pragma unroll
                for (int j = kk + 1; j < 32; j++) a[j] = EXACT ? xmsub(a[j], l, urow[j]) : fmaf(-l, urow[j], a[j]);
            }
        });
        // write the sub-panel: multipliers to L (unit diagonal, zeros above), U11 entries to the working matrix
        if (live) {
            float4* wdst = reinterpret_cast<float4*>(wk + (long)row * n + cs);
            float4* ldst = reinterpret_cast<float4*>(lm + (long)row * n + cs);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float u[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int c = cs + 4 * q + e;
                    const float v = a[4 * q + e];
                    u[e] = (row <= c) ? v : 0.0f;
                    l[e] = (row > c) ? v : (row == c ? 1.0f : 0.0f);
                }
                wdst[q] = make_float4(u[0], u[1], u[2], u[3]);
                ldst[q] = make_float4(l[0], l[1], l[2], l[3]);
            }
            if (row >= cs && row < cs + 32) {
#pragma unroll
                for (int j = 0; j < 32; j++) l11[row - cs][j] = a[j];
            }
        }
        const int rem0 = cs + 32, nrem = o + NB - rem0;  // columns of the panel right of the sub-panel
        __syncthreads();
        // the sub-panel's row exchanges on the panel columns LEFT of it (multipliers already in the L buffer)
        if (t >= 96 && t < 96 + (cs - o)) {
            float* col = lm + o + (t - 96);
            for (int kk = 0; kk < 32; kk++) {
                const int k = cs + kk, p = spiv[kk];
                if (p != k) {
                    const float x = col[(long)k * n], y = col[(long)p * n];
                    col[(long)k * n] = y;
                    col[(long)p * n] = x;
                }
            }
        }
        if (nrem > 0) {
            // 1. the sub-panel's row exchanges on those columns (one thread per column, in pivot order)
            if (t < nrem) {
                float* col = wk + rem0 + t;
                for (int kk = 0; kk < 32; kk++) {
                    const int k = cs + kk, p = spiv[kk];
                    if (p != k) {
                        const float x = col[(long)k * n], y = col[(long)p * n];
                        col[(long)k * n] = y;
                        col[(long)p * n] = x;
                    }
                }
                // 2. U rows of the sub-panel: forward substitution with the unit-lower 32 x 32 block
                float x[32];
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    float s = col[(long)(cs + i) * n];
#pragma unroll
                    for (int kk = 0; kk < i; kk++) s = EXACT ? xmsub(s, l11[i][kk], x[kk]) : fmaf(-l11[i][kk], x[kk], s);
                    x[i] = s;
                    col[(long)(cs + i) * n] = s;
                    ublk[i][t] = s;
                }
            }
            __syncthreads();
            // 3. rank-32 update of the rows below the sub-panel
            if (live && row >= cs + 32) {
                float4* dst = reinterpret_cast<float4*>(wk + (long)row * n + rem0);
                for (int q = 0; q < nrem / 4; q++) {
                    float4 v = dst[q];
#pragma unroll
                    for (int kk = 0; kk < 32; kk++) {
                        const float4 u = *reinterpret_cast<const float4*>(&ublk[kk][4 * q]);
                        if constexpr (EXACT) {
                            v.x = xmsub(v.x, a[kk], u.x), v.y = xmsub(v.y, a[kk], u.y);
                            v.z = xmsub(v.z, a[kk], u.z), v.w = xmsub(v.w, a[kk], u.w);
                        } else {
                            v.x = fmaf(-a[kk], u.x, v.x), v.y = fmaf(-a[kk], u.y, v.y);
                            v.z = fmaf(-a[kk], u.z, v.z), v.w = fmaf(-a[kk], u.w, v.w);
                        }
                    }
                    dst[q] = v;
                }
            }
            __syncthreads();
        }
    }
}

// Slots of the rows a panel's 128 exchanges touch: slots 0..127 are positions o..o+127, further slots are
// the distinct pivot rows below the panel.  pslot[kk] = slot of P[o + kk].  Returns the slot count.
// `slotmap` (n ints of shared memory, set to -1 by the caller) maps a row below the panel to its slot.
__device__ __forceinline__ int lu_touched_rows(const int* piv, int o, int* pos, int* pslot, int* slotmap) {
    int cnt = NB;
    for (int kk = 0; kk < NB; kk++) pos[kk] = o + kk;
    for (int kk = 0; kk < NB; kk++) {
        const int p = piv[kk];
        int sp;
        if (p < o + NB) sp = p - o;
        else {
            sp = slotmap[p];
            if (sp < 0) {
                sp = cnt;
                slotmap[p] = cnt;
                pos[cnt++] = p;
            }
        }
        pslot[kk] = sp;
    }
    return cnt;
}

// The panel's 128 row exchanges on the columns right of the panel (working matrix).  The exchanges are
// first composed into "slot q receives the old row of slot src[q]" (one thread, shared memory), then every
// thread moves its column with independent loads followed by independent stores instead of 128 dependent
// exchange steps.  grid: (ceil(ncols / 128), batch), 128 threads
__global__ void __launch_bounds__(128)
lu_swap_kernel(float* __restrict__ Wk, const int* __restrict__ P, int n, int o) {
    extern __shared__ int slotmap[];  // n ints
    __shared__ int piv[NB], pos[2 * NB], pslot[NB], src[2 * NB];
    __shared__ int ntouch;
    piv[threadIdx.x] = P[(long)blockIdx.y * n + o + threadIdx.x];
    for (int i = threadIdx.x; i < n; i += 128) slotmap[i] = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cnt = lu_touched_rows(piv, o, pos, pslot, slotmap);
        for (int q = 0; q < cnt; q++) src[q] = q;
        for (int kk = 0; kk < NB; kk++) {
            const int sp = pslot[kk], tmp = src[kk];
            src[kk] = src[sp];
            src[sp] = tmp;
        }
        ntouch = cnt;
    }
    __syncthreads();
    const int j = o + NB + blockIdx.x * 128 + threadIdx.x;
    if (j >= n) return;
    float* col = Wk + (long)blockIdx.y * n * n + j;
    const int cnt = ntouch;
    float buf[2 * NB];  // per-thread staging (local memory, L1-resident): all reads precede all writes
    for (int q = 0; q < cnt; q++) buf[q] = col[(long)pos[src[q]] * n];
    for (int q = 0; q < cnt; q++)
        if (src[q] != q) col[(long)pos[q] * n] = buf[q];
}

// Back to the reference's convention for the panel's own 128 L columns.  Inside the panel rows were
// exchanged whole (the deferred block updates need L in final row order); in the reference the multipliers
// of column c stay where they were computed and are not carried along by later exchanges
// (lu_v1.1/model4x4/lu.cpp:55-63 exchanges columns >= id only).  Undo, per column c, the panel's exchanges
// k' > c, last one first; exchanges of later panels never touch these columns.  The touched rows
// (<= 256) x 128 columns are staged in shared memory.  grid: (batch), kUnswapThreads threads
__global__ void __launch_bounds__(1024)
lu_unswap_panel_kernel(float* __restrict__ L, const int* __restrict__ P, int n, int o) {
    extern __shared__ __align__(16) float us_tile[];  // [2 * NB][NB + 1] floats, then n ints (slot map)
    __shared__ int piv[NB], pos[2 * NB], pslot[NB];
    __shared__ int ntouch;
    constexpr int ld = NB + 1;
    int* slotmap = reinterpret_cast<int*>(us_tile + 2 * NB * ld);
    if (threadIdx.x < NB) piv[threadIdx.x] = P[(long)blockIdx.x * n + o + threadIdx.x];
    for (int i = threadIdx.x; i < n; i += blockDim.x) slotmap[i] = -1;
    __syncthreads();
    if (threadIdx.x == 0) ntouch = lu_touched_rows(piv, o, pos, pslot, slotmap);
    __syncthreads();
    const int cnt = ntouch;
    float* base = L + (long)blockIdx.x * n * n + o;
    for (int e = threadIdx.x; e < cnt * NB; e += blockDim.x) {
        const int q = e >> 7, c = e & (NB - 1);
        us_tile[q * ld + c] = base[(long)pos[q] * n + c];
    }
    __syncthreads();
    if (threadIdx.x < NB) {
        const int c = threadIdx.x;  // column o + c: undo exchanges kk = 127 .. c+1
        for (int kk = NB - 1; kk > c; kk--) {
            const int sp = pslot[kk];
            if (sp != kk) {
                const float a = us_tile[kk * ld + c], b = us_tile[sp * ld + c];
                us_tile[kk * ld + c] = b;
                us_tile[sp * ld + c] = a;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * NB; e += blockDim.x) {
        const int q = e >> 7, c = e & (NB - 1);
        base[(long)pos[q] * n + c] = us_tile[q * ld + c];
    }
}

constexpr int kUnswapThreads = 1024;  // the gather / scatter of the touched rows is latency bound: more loads in flight

// W = inv(L11) for the Crout LU (lu_crout.cu): L11 = rows posmap[o .. o+127] of M2, columns o .. o+127 (unit lower)
cudaError_t lu_inverse_l11(const float* M2, const int* posmap, float* W, int n, int o, long batch, cudaStream_t st) {
    cudaError_t e = ensure_dynamic_smem(blk_diag2_kernel<true>, Diag2::kBytes);
    if (e != cudaSuccess) return e;
    blk_diag2_kernel<true><<<(unsigned)batch, 256, Diag2::kBytes, st>>>(M2, W, W, n, o, posmap);
    count_launch();
    return cudaGetLastError();
}

// Blocked LU in the reference's exact arithmetic (HLSD_FLAG_EXACT / HLSD_PATH_BLOCKED, 256 <= n <= 1024): the structure
// of lu_blocked with every arithmetic step exact and in the reference's per-element order - the pivoted panel kernel
// in exact mode, the same exchange / un-exchange kernels (pure data movement), U12 by forward substitution and the
// trailing update on SIMT tiles (blocked_exact.cu).  Bit-identical to the csim: L, U and P.
cudaError_t xlu_u12_and_update(float* Wk, const float* L, int n, int o, long batch, cudaStream_t st);  // blocked_exact.cu

cudaError_t lu_blocked_exact(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, cudaStream_t st) {
    if (n % NB != 0 || n < 2 * NB || n > 1024) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L | (uintptr_t)U) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    if (batch > 65535) return cudaErrorNotSupported;
    cudaError_t e;
    const size_t unswap_smem = (size_t)2 * NB * (NB + 1) * sizeof(float) + (size_t)n * sizeof(int);
    if ((e = ensure_dynamic_smem(lu_unswap_panel_kernel, unswap_smem)) != cudaSuccess) return e;
    const unsigned nb = (unsigned)batch;
    const long nn = (long)n * n;
    {
        ProfScope ps(HLSD_PROF_OTHER, st);
        e = cudaMemcpyAsync(U, A, sizeof(float) * (size_t)batch * nn, cudaMemcpyDeviceToDevice, st);
    }
    if (e != cudaSuccess) return e;
    const int panels = n / NB;
    {
        ProfScope ps(HLSD_PROF_OTHER, st);
        blk_zero_upper_kernel<<<dim3(panels * (panels - 1) / 2, nb), 256, 0, st>>>(L, n);
        count_launch();
    }
    for (int p = 0; p < panels; p++) {
        const int o = p * NB;
        const int ncols = (panels - p - 1) * NB;
        {
            ProfScope ps(HLSD_PROF_PANEL, st);
            lu_panel_kernel<true><<<nb, std::max(256, n - o), 0, st>>>(U, L, P, n, o, pivoting);
            count_launch();
        }
        if (ncols > 0) {
            if (pivoting) {
                ProfScope ps(HLSD_PROF_PERMUTE, st);
                lu_swap_kernel<<<dim3((ncols + 127) / 128, nb), 128, n * sizeof(int), st>>>(U, P, n, o);
                count_launch();
            }
            if ((e = xlu_u12_and_update(U, L, n, o, batch, st)) != cudaSuccess) return e;
        }
        if (pivoting) {
            ProfScope ps(HLSD_PROF_PERMUTE, st);
            lu_unswap_panel_kernel<<<nb, kUnswapThreads, unswap_smem, st>>>(L, P, n, o);  // L21 has served as operand
            count_launch();
        }
    }
    return cudaGetLastError();
}

}  // namespace hlsd
