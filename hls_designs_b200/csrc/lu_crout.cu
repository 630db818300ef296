// Blocked LU with partial pivoting for large N on the tensor cores (N a multiple of 128, 256 <= N <= 1024), batched.
//
// Same elimination as the reference's LU designs (LU/lu_v1.1/model4x4/lu.cpp:23-92: pivot = first row with the
// largest |a_ik|, rows exchanged in columns >= k only, multipliers l_ik = a_ik / a_kk, a_ij -= l_ik u_kj),
// reorganised so that (1) NO ROW IS EVER MOVED and (2) every element of the trailing matrix is produced ONCE:
//
//   * Rows stay where they lie in A.  `posmap` (position in the reference's exchanged order -> physical row) is the
//     only thing a pivot exchange changes; every kernel addresses rows through it.  The operand A is never copied or
//     modified, the multipliers live in a workspace M2[physical row][step] whose row order never changes, and U is
//     kept a second time transposed (UT[c][k] = U[k][c]) because the tensor core wants both operands k-contiguous.
//   * Crout order (left-looking in both directions) over 128-column panels p, o = 128 p:
//       K1 lu_ll_kernel<LL_COL>   (tcgen05)  X[i][c] = A[pm[i]][o+c] - sum_{k<o} M2[pm[i]][k] UT[o+c][k]   positions i >= o
//                                            -> transposed panel workspace PWT[c][i-o]        (panel 0: plain transposing copy)
//       K2 lu_cpanel_kernel       (SIMT)     pivoted elimination of the panel, thread = row, positions tracked in
//                                            registers, ONE barrier per column; writes M2, U11, P, the new posmap
//       K3 blk_diag2_kernel<true> (SIMT)     W = inv(L11), L11 gathered from M2 through posmap
//       K4 lu_ll_kernel<LL_ROW>   (tcgen05)  X12[m][c] = A[piv_m][c] - sum_{k<o} M2[piv_m][k] UT[c][k]  (the 128 pivot rows), stored
//                                            transposed (T0[c][m])
//          lu_trsm_kernel         (tcgen05)  U12^T = X12^T W^T  -> UT[c][o..o+127] and U[o..o+127][c]
//     so C is read once (from A) and written once per element, and the K loops are 128 p long (as in the Cholesky
//     path's blk_ll2_kernel) instead of a rank-128 update per panel with a read-modify-write of the trailing matrix.
//   * L in the reference's LINPACK layout (the multiplier of step k stays at the POSITION its row had at step k) is
//     assembled from M2 at the end by lu_assemble_l_kernel, which replays each panel's exchanges on the <= 256
//     positions they touch.
// The data flow is modelled in scripts/lu_crout_model.py and checked against a plain LINPACK elimination on the CPU
// (tests/test_lu_crout_model.py).  Tensor-core arithmetic is 3xTF32 with fp32 accumulation (see blocked.cu): not
// bit-identical to the csim; tolerances in tests/test_gpu_tensor.py.  HLSD_FLAG_EXACT selects lu_blocked_exact.
#include "kernels.h"
#include "tc.cuh"

namespace hlsd {

// ------------------------------------------------------------------------------------------------
// Zero fill of the blocks strictly above the block diagonal of L and strictly below that of U, and the identity
// posmap of panel 0.  grid: (panels (panels - 1) / 2, batch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lu_zero_blocks_kernel(float* __restrict__ L, float* __restrict__ U, int* __restrict__ hist0, int n) {
    int r = blockIdx.x;  // unrank into (bi < bj)
    int bj = (int)((sqrtf(8.0f * r + 1.0f) + 1.0f) * 0.5f);
    while (bj * (bj - 1) / 2 > r) bj--;
    while ((bj + 1) * bj / 2 <= r) bj++;
    const int bi = r - bj * (bj - 1) / 2;
    float4* lb = reinterpret_cast<float4*>(L + (long)blockIdx.y * n * n + (long)bi * NB * n + bj * NB);
    float4* ub = reinterpret_cast<float4*>(U + (long)blockIdx.y * n * n + (long)bj * NB * n + bi * NB);
    const int n4 = n / 4;
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int e = threadIdx.x; e < NB * (NB / 4); e += 256) {
        const int row = e / (NB / 4), c4 = e % (NB / 4);
        __stcs(lb + (long)row * n4 + c4, z);
        __stcs(ub + (long)row * n4 + c4, z);
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n; i += 256) hist0[(long)blockIdx.y * n + i] = i;
}

// Panel 0: PWT[c][i] = A[i][c], c < 128 (no earlier panels to subtract).  grid: (n / 32, 4, batch), block (32, 8)
__global__ void __launch_bounds__(256)
lu_panel0_load_kernel(const float* __restrict__ A, float* __restrict__ PWT, int n) {
    __shared__ float tile[32][33];
    const float* a = A + (long)blockIdx.z * n * n;
    float* pw = PWT + (long)blockIdx.z * n * NB;
    const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = a[(long)(i0 + r) * n + c0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) pw[(long)(c0 + r) * n + i0 + tx] = tile[tx][r];
}

// ------------------------------------------------------------------------------------------------
// Long-K tensor-core kernel of the Crout scheme: a product over K = o (all finished panels at once), result =
// (tile of A, through posmap) - product.  The pipeline is blk_ll2_kernel's (blocked.cu): 512 threads load fp32 operand
// chunks three chunks ahead into registers, write each chunk to shared memory as the raw (= tf32 hi) and the lo tile in
// the canonical K-major SWIZZLE_128B layout, one thread issues the 3xTF32 MMAs into TMEM.  A CTA produces two paired
// tiles that share one operand: the SHARED tile (128 rows) is the M operand, the two paired tiles - contiguous in shared
// memory - form one 256-row N operand, so the accumulator is 128 lanes (rows of the shared tile) x 256 columns (rows of
// the paired tiles), i.e. the result comes out transposed and the epilogue thread = a row of the shared tile.
//   LL_COL (K1): paired tiles = positions o + 128 (2 bx + h) + r, rows M2[pm[position]]; shared tile = the panel's
//                columns c, rows UT[o + c].  Thread = column c: the tile of A is read coalesced over c, the result goes to
//                PWT[c][position - o], 32 consecutive positions per thread.
//   LL_ROW (K4): paired tiles = columns c = o + 128 + 128 (2 bx + h) + r of U12, rows UT[c]; shared tile = the panel's
//                128 pivot rows, rows M2[pm[o + m]] (pm = posmap AFTER the panel).  Thread = pivot row m: the tile of A is
//                read from the thread's own row, the result X12^T goes to T0[c - o - 128][m], coalesced over m.
// grid: (ceil(tiles / 2), batch)
// ------------------------------------------------------------------------------------------------
struct LlSmem {
    static constexpr int kTile = 128 * 32 * 4;
    static constexpr int kRaw = 0;           // [2 stages][A0, A1, B]
    static constexpr int kLo = 6 * kTile;    // [2 stages][A0, A1, B]
    static constexpr int kBar = 12 * kTile;
    static constexpr int kTmemPtr = kBar + 16;
    static constexpr int kTotal = kBar + 32 + 1024;
};
enum { LL_COL = 0, LL_ROW = 1 };

template <int MODE>
__global__ void __launch_bounds__(512, 1)
lu_ll_kernel(const float* __restrict__ A, const float* __restrict__ M2, const float* __restrict__ UT,
             const int* __restrict__ posmap, float* __restrict__ out, int n, int o) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LlSmem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + LlSmem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long b = blockIdx.y;
    const float* Ab = A + b * n * n;
    const float* M2b = M2 + b * n * n;
    const float* UTb = UT + b * n * n;
    const int* pm = posmap + b * n;
    const int tiles = MODE == LL_COL ? (n - o) / NB : (n - o - NB) / NB;
    const int t0 = 2 * blockIdx.x;
    const bool two = t0 + 1 < tiles;
    const int C = o / 32;  // K chunks (a multiple of 4)

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

    // operand rows of this thread's two 16-byte pieces per tile (rows t/8 and 64 + t/8, bytes 16 (t%8) of each chunk)
    const float *pA0[2], *pA1[2], *pB[2];
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int row = it * 64 + (t >> 3), c4 = t & 7;
        if constexpr (MODE == LL_COL) {
            pA0[it] = M2b + (long)pm[o + t0 * NB + row] * n + c4 * 4;
            pA1[it] = two ? M2b + (long)pm[o + (t0 + 1) * NB + row] * n + c4 * 4 : pA0[it];
            pB[it] = UTb + (long)(o + row) * n + c4 * 4;
        } else {
            pA0[it] = UTb + (long)(o + NB + t0 * NB + row) * n + c4 * 4;
            pA1[it] = two ? pA0[it] + (long)NB * n : pA0[it];
            pB[it] = M2b + (long)pm[o + row] * n + c4 * 4;
        }
    }
    auto piece_off = [&](int it) {
        const int row = it * 64 + (t >> 3), c4 = t & 7;
        return row * 128 + ((c4 ^ (row & 7)) << 4);
    };
    constexpr int kPre = 3;  // chunks of global loads in flight per thread (register stages)
    float4 b0[kPre][2], b1[kPre][2], bb[kPre][2];  // [register stage][piece]: A0, A1, B
    auto gload = [&](int c, int sreg) {
        if (c < C) {
#pragma unroll
            for (int it = 0; it < 2; it++) {
                b0[sreg][it] = __ldg(reinterpret_cast<const float4*>(pA0[it] + c * 32));
                if (two) b1[sreg][it] = __ldg(reinterpret_cast<const float4*>(pA1[it] + c * 32));
                bb[sreg][it] = __ldg(reinterpret_cast<const float4*>(pB[it] + c * 32));
            }
        }
    };
    auto step = [&](int c, int sreg) {
        if (c >= 2) {
            mbar_wait(&bars[c & 1], (uint32_t)(((c - 2) >> 1) & 1));
            tc_fence_after();
        }
        unsigned char* ra0 = smem + LlSmem::kRaw + (c & 1) * 3 * LlSmem::kTile;
        unsigned char* la0 = smem + LlSmem::kLo + (c & 1) * 3 * LlSmem::kTile;
        unsigned char* ra1 = ra0 + LlSmem::kTile;
        unsigned char* la1 = la0 + LlSmem::kTile;
        unsigned char* rb = ra0 + 2 * LlSmem::kTile;
        unsigned char* lb = la0 + 2 * LlSmem::kTile;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int off = piece_off(it);
            *reinterpret_cast<float4*>(ra0 + off) = b0[sreg][it];
            *reinterpret_cast<float4*>(la0 + off) = lo_of(b0[sreg][it]);
            if (two) {
                *reinterpret_cast<float4*>(ra1 + off) = b1[sreg][it];
                *reinterpret_cast<float4*>(la1 + off) = lo_of(b1[sreg][it]);
            }
            *reinterpret_cast<float4*>(rb + off) = bb[sreg][it];
            *reinterpret_cast<float4*>(lb + off) = lo_of(bb[sreg][it]);
        }
        gload(c + kPre, sreg);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (t == 0) {
            // ONE accumulator of 128 x 256: the shared tile is the M operand, the two paired tiles - contiguous in
            // shared memory - form one 256-row N operand, so the shared tile is read once per k-step instead of twice
            // (the kernel is bound by shared-memory bandwidth: 12 KB instead of 16 KB of operand reads per 2 x 262 kflop)
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
    // three chunks of loads in flight (ncu: with two, `long_scoreboard` was the top stall and DRAM and the tensor pipe
    // both sat near 50 %): the register stage is c % 3 (static), the shared-memory stage c % 2
    gload(0, 0);
    gload(1, 1);
    gload(2, 2);
    for (int c = 0; c < C; c += 3) {
        step(c, 0);
        if (c + 1 < C) step(c + 1, 1);
        if (c + 2 < C) step(c + 2, 2);
    }
    // epilogue: thread = accumulator lane drow = row of the SHARED operand tile; accumulator columns h * 128 + col .. + 31 =
    // rows col .. col + 31 of paired tile h.  The tile of A is fetched while the last chunks are still being multiplied.
    const int drow = (warp & 3) * 32 + lane;
    const int col = (warp >> 2) * 32;
    float cv[2][32];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (h == 1 && !two) break;
        if constexpr (MODE == LL_COL) {
            // lane = column c of the panel, accumulator column j = position (t0 + h) * 128 + col + j: coalesced over c
            const float* src = Ab + o + drow;
            const int* pp = pm + o + (t0 + h) * NB + col;
#pragma unroll
            for (int j = 0; j < 32; j++) cv[h][j] = __ldg(src + (long)pp[j] * n);
        } else {
            // lane = pivot row m, accumulator column j = matrix column o + 128 + (t0 + h) * 128 + col + j
            const float4* src = reinterpret_cast<const float4*>(Ab + (long)pm[o + drow] * n + o + NB + (t0 + h) * NB + col);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 v = __ldg(src + q);
                cv[h][4 * q] = v.x, cv[h][4 * q + 1] = v.y, cv[h][4 * q + 2] = v.z, cv[h][4 * q + 3] = v.w;
            }
        }
    }
    if (C > 0) {
        mbar_wait(&bars[(C - 2) & 1], (uint32_t)(((C - 2) >> 1) & 1));
        mbar_wait(&bars[(C - 1) & 1], (uint32_t)(((C - 1) >> 1) & 1));
        tc_fence_after();
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (h == 1 && !two) break;
        float acc[32];
        if (C > 0) {
            tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * 128 + col), acc);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) acc[j] = 0.0f;
        }
        if constexpr (MODE == LL_COL) {
            // PWT[c][position - o]: 32 consecutive positions per thread
            float4* dst = reinterpret_cast<float4*>(out + b * n * NB + (long)drow * n + (t0 + h) * NB + col);
#pragma unroll
            for (int q = 0; q < 8; q++)
                dst[q] = make_float4(cv[h][4 * q] - acc[4 * q], cv[h][4 * q + 1] - acc[4 * q + 1],
                                     cv[h][4 * q + 2] - acc[4 * q + 2], cv[h][4 * q + 3] - acc[4 * q + 3]);
        } else {
            // T0[c][m]: coalesced over m
            float* dst = out + b * n * NB + (long)((t0 + h) * NB + col) * NB + drow;
#pragma unroll
            for (int j = 0; j < 32; j++) dst[(long)j * NB] = cv[h][j] - acc[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// ------------------------------------------------------------------------------------------------
// U12^T = X12^T W^T: one 128 x 128 x 128 tile per CTA (the pipeline of blk_mma_kernel), A operand = T0 rows (columns c
// of the matrix), B operand = W = inv(L11).  The result goes out twice: UT[c][o + k'] (row-wise, through a swizzled
// shared-memory stage) and U[o + k'][c] (straight from the accumulator registers: thread = c, coalesced over c).
// grid: (column blocks right of the panel, batch)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, MmaSmem::kStages == 1 ? 3 : 2)
lu_trsm_kernel(const float* __restrict__ T0, const float* __restrict__ W, float* __restrict__ UT, float* __restrict__ U,
               int n, int o) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MmaSmem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MmaSmem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long b = blockIdx.y;
    const int cb = blockIdx.x;
    const float* Arows = T0 + b * n * NB + (long)cb * NB * NB;
    const float* Brows = W + b * NB * NB;
    if (warp == 0) tmem_alloc(tmem_ptr, 128);
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;
    constexpr int kChunks = NB / MmaSmem::kChunkK;
    auto piece_off = [&](int it) {
        const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
        return row * 128 + ((c4 ^ (row & 7)) << 4);
    };
    unsigned char* ra = smem + MmaSmem::kRawA;
    unsigned char* rb = smem + MmaSmem::kRawB;
    unsigned char* la = smem + MmaSmem::kLoA;
    unsigned char* lb = smem + MmaSmem::kLoB;
    auto issue_loads = [&](int chunk) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
            cp_async16(ra + piece_off(it), Arows + row * NB + chunk * MmaSmem::kChunkK + c4 * 4);
            cp_async16(rb + piece_off(it), Brows + row * NB + chunk * MmaSmem::kChunkK + c4 * 4);
        }
        cp_async_commit();
    };
    issue_loads(0);
    for (int chunk = 0; chunk < kChunks; chunk++) {
        cp_async_wait<0>();
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int off = piece_off(it);
            *reinterpret_cast<float4*>(la + off) = lo_of(*reinterpret_cast<const float4*>(ra + off));
            *reinterpret_cast<float4*>(lb + off) = lo_of(*reinterpret_cast<const float4*>(rb + off));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (t == 0) {
            const uint64_t dAhi = make_sw128_desc(smem_u32(ra)), dAlo = make_sw128_desc(smem_u32(la));
            const uint64_t dBhi = make_sw128_desc(smem_u32(rb)), dBlo = make_sw128_desc(smem_u32(lb));
#pragma unroll
            for (int kk = 0; kk < MmaSmem::kChunkK / 8; kk++) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t first = (chunk == 0 && kk == 0) ? 0u : 1u;
                umma_tf32(tmem_d, dAlo + adv, dBhi + adv, kIdescTf32, first);
                umma_tf32(tmem_d, dAhi + adv, dBlo + adv, kIdescTf32, 1u);
                umma_tf32(tmem_d, dAhi + adv, dBhi + adv, kIdescTf32, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, (uint32_t)(chunk & 1));
        tc_fence_after();
        if (chunk + 1 < kChunks) issue_loads(chunk + 1);
    }
    {
        float* stage = reinterpret_cast<float*>(smem);
        const int drow = (warp & 3) * 32 + lane;  // column c of the matrix (relative to the column block)
        const int dcol0 = (warp >> 2) * 64;       // k' range
        float* ucol = U + b * n * n + (long)o * n + o + NB + cb * NB + drow;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float acc[32];
            const int col = dcol0 + h * 32;
            tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, acc);
#pragma unroll
            for (int j = 0; j < 32; j++) __stcs(ucol + (long)(col + j) * n, acc[j]);  // U[o + k'][c]
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int c4 = (col >> 2) + q;
                *reinterpret_cast<float4*>(stage + drow * 128 + ((c4 ^ (drow & 7)) << 2)) =
                    make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        __syncthreads();
        float* utb = UT + b * n * n + (long)(o + NB + cb * NB) * n + o;
#pragma unroll 4
        for (int rr = 0; rr < 16; rr++) {
            const int row = warp * 16 + rr;
            const float4 d = *reinterpret_cast<const float4*>(stage + row * 128 + ((lane ^ (row & 7)) << 2));
            *(reinterpret_cast<float4*>(utb + (long)row * n) + lane) = d;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 128);
}

// ------------------------------------------------------------------------------------------------
// Panel factorisation: one CTA per matrix, thread t = the row at position o + t when the panel starts (physical row
// pm_in[o + t]); the row's 32 values of the current sub-panel live in registers, the rest of the panel in the
// transposed workspace PWT[c][t] (every access coalesced over t).  Rows never move: a thread only tracks its position
// (`mypos`); the thread that wins the pivot search of step k takes position k and retires, the thread that sat at
// position k takes the winner's old position - exactly the reference's exchange (lu.cpp:44-63), on indices.
// Pivot search with ONE barrier per column: the key (|a| as an unsigned integer, +1; 0 = not a candidate) is
// maximised over the warp with a single REDUX, ties go to the smallest position with a second one (the reference's
// downward scan with a strict '<'); each warp's winner publishes (key, position) and - speculatively - its row into
// buffers double-buffered by the parity of the column; after the barrier every warp reduces the <= 32 warp candidates
// the same way and reads the winning row.  NaN handling as in the exact kernels (lu.cu: a NaN never wins unless it sits
// at position k, where it is never displaced).
// Between sub-panels: the 32 new pivot rows' values in the remaining panel columns are gathered, U12 of the sub-panel
// follows by forward substitution with the unit-lower block held in `prow`, and every row that is not a pivot yet
// subtracts its 32 multipliers times those U rows from its part of PWT.
// Outputs: M2[physical row][o .. o+127] (multipliers; for pivot rows, right of their own step, U values that nobody
// reads), U[o+kk][o .. o+127] (zeros left of the diagonal), P, and the posmap after the panel's exchanges.
// ------------------------------------------------------------------------------------------------
struct PCand {
    unsigned key;
    int pos;
};
constexpr int kXposeWarpBytes = 32 * 36 * 4;  // dynamic shared memory of lu_cpanel_kernel per warp (sub-panel width 32)

// RPT = rows per thread (thread t holds the rows at positions o + t and o + t + blockDim.x): the per-column work that
// does not depend on the row count - key reduction, publication, the barrier, the winner test - is paid once per
// thread, so two rows per thread nearly halve the instructions of the column loop, and the U rows read from shared
// memory in the rank-32 update serve eight instead of four multiply-adds.
template <int RPT, int SW>
__global__ void __launch_bounds__(1024 / RPT, SW == 16 ? 2 : 1)
lu_cpanel_kernel(float* __restrict__ PWT, float* __restrict__ M2, float* __restrict__ U, int* __restrict__ P,
                 const int* __restrict__ pm_in, int* __restrict__ pm_out, int n, int o, int pivoting) {
    __shared__ __align__(16) float slot[2][32][SW];  // speculative pivot rows, one per warp
    __shared__ PCand cand[2][32];
    __shared__ __align__(16) float prow[SW][SW + 4];     // pivot rows of the sub-panel: multipliers | U11
    __shared__ int tw[SW];                           // row (index into PWT) that won step kk
    __shared__ __align__(16) float ublk[SW][NB - SW + 4];  // the sub-panel's U rows in the remaining panel columns
    extern __shared__ __align__(16) float xpose[];   // [warps][32][SW + 4]: per-warp transposition stage of the M2 stores
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int T = blockDim.x, nwarps = T >> 5;
    const int R = n - o;
    const long b = blockIdx.x;
    float* pwb = PWT + b * n * NB;  // element (c, row): pwb[c * n + row]
    int phys[RPT], mypos[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int ri = t + i * T;
        phys[i] = ri < R ? pm_in[b * n + o + ri] : 0;
        mypos[i] = ri < R ? o + ri : -1;
    }
    constexpr unsigned kFull = 0xffffffffu;
    for (int sp = 0; sp < NB / SW; sp++) {
        const int cs = SW * sp;
        float a[RPT][SW];
#pragma unroll
        for (int i = 0; i < RPT; i++)
#pragma unroll
            for (int j = 0; j < SW; j++) a[i][j] = t + i * T < R ? pwb[(long)(cs + j) * n + t + i * T] : 0.0f;
        static_for<0, SW>([&](auto kkc) {
            constexpr int kk = decltype(kkc)::value;
            constexpr int par = kk & 1;
            const int k = o + cs + kk;
            unsigned kb = 0;
            int pb = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                unsigned key = 0;
                if (mypos[i] >= k) {
                    const float v = fabsf(a[i][kk]);
                    key = (v != v) ? (mypos[i] == k ? 0x7f800001u : 0u) : __float_as_uint(v) + 1u;
                    if (!pivoting) key = mypos[i] == k ? 1u : 0u;  // HLSD_LU_NOPIVOT: the row at position k is the only candidate
                }
                if (key > kb || (key == kb && key != 0u && mypos[i] < pb)) kb = key, pb = mypos[i];
            }
            const unsigned wmax = __reduce_max_sync(kFull, kb);
            const unsigned cpos = (kb == wmax && wmax != 0u) ? (unsigned)pb : 0x7fffffffu;
            const unsigned wpos = __reduce_min_sync(kFull, cpos);
            if (lane == 0) cand[par][warp] = PCand{wmax, (int)wpos};  // (no candidate: key 0, position 0x7fffffff)
            if (cpos == wpos && wmax != 0u) {
                float4* dst = reinterpret_cast<float4*>(slot[par][warp]);
                if (RPT == 2 && mypos[RPT - 1] == (int)wpos) {
#pragma unroll
                    for (int q = kk / 4; q < SW / 4; q++)
                        dst[q] = make_float4(a[RPT - 1][4 * q], a[RPT - 1][4 * q + 1], a[RPT - 1][4 * q + 2], a[RPT - 1][4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = kk / 4; q < SW / 4; q++) dst[q] = make_float4(a[0][4 * q], a[0][4 * q + 1], a[0][4 * q + 2], a[0][4 * q + 3]);
                }
            }
            __syncthreads();
            const PCand c = lane < nwarps ? cand[par][lane] : PCand{0u, 0x7fffffff};
            const unsigned gmax = __reduce_max_sync(kFull, c.key);
            const unsigned gcp = (c.key == gmax) ? (unsigned)c.pos : 0x7fffffffu;
            const int gp = (int)__reduce_min_sync(kFull, gcp);
            const int wwin = __ffs(__ballot_sync(kFull, c.key == gmax && c.pos == gp)) - 1;
            const float* urow = slot[par][wwin];
            if (t == 0) P[b * n + k] = gp;
            bool act[RPT];
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                if (mypos[i] == gp) {  // this row is the pivot row: it takes position k and retires
                    tw[kk] = t + i * T;
                    float4* dst = reinterpret_cast<float4*>(prow[kk]);
#pragma unroll
                    for (int q = 0; q < SW / 4; q++) dst[q] = make_float4(a[i][4 * q], a[i][4 * q + 1], a[i][4 * q + 2], a[i][4 * q + 3]);
                    mypos[i] = k;
                } else if (mypos[i] == k) {
                    mypos[i] = gp;
                }
                act[i] = mypos[i] > k;
            }
            bool any = act[0];
#pragma unroll
            for (int i = 1; i < RPT; i++) any |= act[i];
            if (any) {
                // quotient = reciprocal (MUFU) times numerator, corrected once with the exact residual
                const float piv = urow[kk];
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(piv));
                float l[RPT];
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    const float l0 = a[i][kk] * r;
                    l[i] = fmaf(fmaf(-piv, l0, a[i][kk]), r, l0);
                    if (act[i]) a[i][kk] = l[i];
                }
                const float4* u4 = reinterpret_cast<const float4*>(urow);
                static_for<(kk + 1) / 4, SW / 4>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    const float4 u = u4[q];
#pragma unroll
                    for (int i = 0; i < RPT; i++) {
                        if (act[i]) {
                            if constexpr (4 * q > kk) a[i][4 * q] = fmaf(-l[i], u.x, a[i][4 * q]);
                            if constexpr (4 * q + 1 > kk) a[i][4 * q + 1] = fmaf(-l[i], u.y, a[i][4 * q + 1]);
                            if constexpr (4 * q + 2 > kk) a[i][4 * q + 2] = fmaf(-l[i], u.z, a[i][4 * q + 2]);
                            if constexpr (4 * q + 3 > kk) a[i][4 * q + 3] = fmaf(-l[i], u.w, a[i][4 * q + 3]);
                        }
                    }
                });
            }
        });
        {
            // M2[physical row][o + cs .. + 31] = the row's 32 registers: transposed through shared memory so that one
            // store instruction writes four complete 128-byte rows instead of a 16-byte piece of thirty-two
            constexpr int XL = SW + 4, PC = SW / 4, RPI = 32 / PC;  // row stride, 16-byte pieces per row, rows per instruction
            float* xp = xpose + warp * (32 * XL);
            float4* mine = reinterpret_cast<float4*>(xp + lane * XL);
            float* m2b = M2 + b * n * n + o + cs + 4 * (lane % PC);
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                if (i > 0) __syncwarp();
#pragma unroll
                for (int q = 0; q < PC; q++) mine[q] = make_float4(a[i][4 * q], a[i][4 * q + 1], a[i][4 * q + 2], a[i][4 * q + 3]);
                __syncwarp();
#pragma unroll
                for (int g = 0; g < 32 / RPI; g++) {
                    const int row = RPI * g + lane / PC;
                    const int row_phys = __shfl_sync(kFull, phys[i], row);
                    const float4 v = *reinterpret_cast<const float4*>(xp + row * XL + 4 * (lane % PC));
                    if (warp * 32 + row + i * T < R) *reinterpret_cast<float4*>(m2b + (long)row_phys * n) = v;
                }
            }
        }
        __syncthreads();  // prow and tw are complete
        const int rem0 = cs + SW, nrem = NB - rem0;  // panel columns right of the sub-panel
        if (nrem > 0) {
            // the new pivot rows in those columns (as they are now: updated by the earlier sub-panels)
            {
                const int gk = lane % SW;   // e % SW is the same for every element e of this thread (blockDim % 32 == 0)
                const int row_t = tw[gk];
                for (int e0 = t; e0 < SW * nrem; e0 += 4 * T) {
                    float g[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int e = e0 + i * T;
                        if (e < SW * nrem) g[i] = pwb[(long)(rem0 + e / SW) * n + row_t];
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int e = e0 + i * T;
                        if (e < SW * nrem) ublk[gk][e / SW] = g[i];
                    }
                }
            }
            __syncthreads();
            // forward substitution with the unit-lower 32 x 32 block (multipliers of the pivot rows): a warp per group of
            // three columns, lane j = row j of the block with its multipliers in registers; row kk of the result travels by
            // shuffle, rows below it subtract their multiple of it (right-looking: 32 steps of shuffle + multiply-add)
            {
                float lrow[SW];
                const int lrow_i = lane < SW ? lane : 0;
                {
                    const float4* src = reinterpret_cast<const float4*>(prow[lrow_i]);
#pragma unroll
                    for (int q = 0; q < SW / 4; q++) {
                        const float4 v = src[q];
                        lrow[4 * q] = v.x, lrow[4 * q + 1] = v.y, lrow[4 * q + 2] = v.z, lrow[4 * q + 3] = v.w;
                    }
                }
                for (int c0 = 3 * warp; c0 < nrem; c0 += 3 * nwarps) {
                    float x[3];
#pragma unroll
                    for (int e = 0; e < 3; e++) x[e] = (c0 + e < nrem && lane < SW) ? ublk[lrow_i][c0 + e] : 0.0f;
#pragma unroll
                    for (int kk = 0; kk < SW - 1; kk++) {
#pragma unroll
                        for (int e = 0; e < 3; e++) {
                            const float xk = __shfl_sync(kFull, x[e], kk);
                            if (lane > kk) x[e] = fmaf(-lrow[kk], xk, x[e]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 3; e++)
                        if (c0 + e < nrem && lane < SW) ublk[lane][c0 + e] = x[e];
                }
            }
            __syncthreads();
            // rank-32 update of the rows that are not pivot rows yet (a thread's retired row is computed along and not stored)
            bool upd[RPT];
            bool any = false;
#pragma unroll
            for (int i = 0; i < RPT; i++) upd[i] = mypos[i] >= o + rem0, any |= upd[i];
            if (any) {
                float v[RPT][4], nx[RPT][4];
                {
                    const float* p0 = pwb + (long)rem0 * n + t;
#pragma unroll
                    for (int i = 0; i < RPT; i++)
#pragma unroll
                        for (int e = 0; e < 4; e++) nx[i][e] = p0[(long)e * n + i * T];
                }
                for (int c4 = 0; c4 < nrem / 4; c4++) {
                    float* p0 = pwb + (long)(rem0 + 4 * c4) * n + t;
#pragma unroll
                    for (int i = 0; i < RPT; i++)
#pragma unroll
                        for (int e = 0; e < 4; e++) v[i][e] = nx[i][e];
                    if (c4 + 1 < nrem / 4) {  // the next four columns' values, in flight during this group's arithmetic
#pragma unroll
                        for (int i = 0; i < RPT; i++)
#pragma unroll
                            for (int e = 0; e < 4; e++) nx[i][e] = p0[(long)(4 + e) * n + i * T];
                    }
#pragma unroll
                    for (int kk = 0; kk < SW; kk++) {
                        const float4 u = *reinterpret_cast<const float4*>(&ublk[kk][4 * c4]);
#pragma unroll
                        for (int i = 0; i < RPT; i++) {
                            v[i][0] = fmaf(-a[i][kk], u.x, v[i][0]), v[i][1] = fmaf(-a[i][kk], u.y, v[i][1]);
                            v[i][2] = fmaf(-a[i][kk], u.z, v[i][2]), v[i][3] = fmaf(-a[i][kk], u.w, v[i][3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < RPT; i++)
                        if (upd[i]) {
#pragma unroll
                            for (int e = 0; e < 4; e++) p0[(long)e * n + i * T] = v[i][e];
                        }
                }
            }
        }
        // rows o+cs .. o+cs+31 of U inside the panel: zeros | U11 of the sub-panel | its U12
        for (int e = t; e < SW * 32; e += T) {
            const int kk = e >> 5, q = e & 31;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = 4 * q + i;
                v[i] = c < cs + kk ? 0.0f : (c < rem0 ? prow[kk][c - cs] : ublk[kk][c - rem0]);
            }
            *reinterpret_cast<float4*>(U + b * n * n + (long)(o + cs + kk) * n + o + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();  // prow, tw, ublk are reused by the next sub-panel
    }
#pragma unroll
    for (int i = 0; i < RPT; i++)
        if (t + i * T < R) pm_out[b * n + mypos[i]] = phys[i];
}

// ------------------------------------------------------------------------------------------------
// L in the reference's layout from M2: L[i][k] = multiplier, at step k, of the row that sat at position i right after
// step k's exchange (lu_v1.1/model4x4/lu.cpp:55-63 exchanges columns >= id only, so a multiplier stays at the position
// where it was computed).  Per (panel, matrix): positions the panel's 128 exchanges never touch keep their occupant
// for the whole panel and are copied row by row.  For the touched positions ("slots": o..o+127 and the distinct pivots
// below, <= 256) one thread replays the exchanges on an occupancy vector and records, per step j, which row was put
// into the pivot's slot q_j; a slot's occupant at column j is then the value of the last such event <= j that names
// it (a ballot and a shuffle per 32 columns), and every output element is read straight from that row of M2 - mostly
// one row per output row, so the loads stay coalesced.  No staging tile: ~6 KB of shared memory, many CTAs per SM.
// grid: (panels, batch), 256 threads
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lu_assemble_l_kernel(const float* __restrict__ M2, float* __restrict__ L, const int* __restrict__ P,
                     const int* __restrict__ hist, int n, int panels) {
    __shared__ int piv[NB], qslot[NB], spos[2 * NB], srow[2 * NB];
    __shared__ short slot_of[1024];
    __shared__ unsigned char occ[2 * NB], evval[NB];
    __shared__ int s_cnt;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int p = blockIdx.x, o = p * NB;
    const long b = blockIdx.y;
    const int* pm = hist + ((long)p * gridDim.y + b) * n;  // posmap at the start of the panel ([panel][batch][n])
    const float* m2 = M2 + b * n * n + o;
    float* lo = L + b * n * n + o;
    if (t < NB) piv[t] = P[b * n + o + t];
    for (int i = t; i < n; i += 256) slot_of[i] = -1;
    __syncthreads();
    if (t < NB && piv[t] >= o + NB) slot_of[piv[t]] = -2;  // marked
    __syncthreads();
    if (warp == 0) {
        // slots 0..127 = positions o..o+127; the marked positions below follow in ascending order
        int base = NB;
        for (int i0 = o + NB; i0 < n; i0 += 32) {
            const bool m = slot_of[i0 + lane] == -2;
            const unsigned bal = __ballot_sync(0xffffffffu, m);
            if (m) {
                const int s = base + __popc(bal & ((1u << lane) - 1u));
                slot_of[i0 + lane] = (short)s;
                spos[s] = i0 + lane;
            }
            base += __popc(bal);
        }
        if (lane == 0) s_cnt = base;
        for (int s = lane; s < NB; s += 32) spos[s] = o + s;
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (t < NB) qslot[t] = piv[t] < o + NB ? piv[t] - o : slot_of[piv[t]];
    for (int s = t; s < cnt; s += 256) {
        occ[s] = (unsigned char)s;
        srow[s] = pm[spos[s]];  // physical row sitting at slot s when the panel starts
    }
    __syncthreads();
    if (t == 0) {
        // replay: step j exchanges the occupants of slot j and slot q_j; evval[j] = what q_j receives
        for (int j = 0; j < NB; j++) {
            const int q = qslot[j];
            const unsigned char x = occ[j], y = occ[q];
            occ[j] = y;
            occ[q] = x;
            evval[j] = x;
        }
    } else if (warp > 0) {
        // meanwhile: positions the panel never touches keep their row for all 128 steps
        for (int i0 = o + NB + (warp - 1) * 4; i0 < n; i0 += 28) {
            float4 v[4];
            bool w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w[i] = i0 + i < n && slot_of[i0 + i] == -1;
                if (w[i]) v[i] = __ldg(reinterpret_cast<const float4*>(m2 + (long)pm[i0 + i] * n) + lane);
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (w[i]) __stcs(reinterpret_cast<float4*>(lo + (long)(i0 + i) * n) + lane, v[i]);
        }
    }
    __syncthreads();
    for (int s = warp; s < cnt; s += 8) {
        const int pos = spos[s];
        int carry = s;  // occupant before the first event: the row that sat here when the panel started
        float v[4];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int j = 32 * g + lane;
            const bool flag = qslot[j] == s;
            const int val = evval[j];
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            const unsigned m = bal & (0xffffffffu >> (31 - lane));  // events at columns <= j of this group
            const int from = __shfl_sync(0xffffffffu, val, m ? 31 - __clz(m) : 0);
            const int here = m ? from : carry;
            if (bal) carry = __shfl_sync(0xffffffffu, val, 31 - __clz(bal));
            const int k = o + j;
            v[g] = pos > k ? __ldg(m2 + (long)srow[here] * n + j) : (pos == k ? 1.0f : 0.0f);
        }
#pragma unroll
        for (int g = 0; g < 4; g++) __stcs(lo + (long)pos * n + 32 * g + lane, v[g]);
    }
}

// ------------------------------------------------------------------------------------------------
// W = inv(L11) through posmap: blocked.cu
cudaError_t lu_inverse_l11(const float* M2, const int* posmap, float* W, int n, int o, long batch, cudaStream_t st);

static cudaError_t lu_crout_slice(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, float* ws, cudaStream_t st) {
    const long nn = (long)n * n;
    const int panels = n / NB;
    float* M2 = ws;
    float* UT = M2 + batch * nn;
    float* PWT = UT + batch * nn;
    float* T0 = PWT + batch * n * NB;
    float* W = T0 + batch * n * NB;
    int* hist = reinterpret_cast<int*>(W + batch * NB * NB);
    const unsigned nb = (unsigned)batch;
    // posmap history, panel-major ([panel][batch][n]): the kernels index a panel's posmaps as base + b * n
    auto pm = [&](int p) { return hist + (long)p * batch * n; };
    {
        ProfScope ps(HLSD_PROF_OTHER, st);
        lu_zero_blocks_kernel<<<dim3(panels * (panels - 1) / 2, nb), 256, 0, st>>>(L, U, pm(0), n);
        count_launch();
        lu_panel0_load_kernel<<<dim3(n / 32, 4, nb), dim3(32, 8), 0, st>>>(A, PWT, n);
        count_launch();
    }
    for (int p = 0; p < panels; p++) {
        const int o = p * NB;
        const int rt = panels - p;       // row tiles of the panel
        const int ct = panels - p - 1;   // column tiles right of it
        if (p > 0) {
            ProfScope ps(HLSD_PROF_UPDATE, st);
            lu_ll_kernel<LL_COL><<<dim3((rt + 1) / 2, nb), 512, LlSmem::kTotal, st>>>(A, M2, UT, pm(p), PWT, n, o);
            count_launch();
        }
        {
            ProfScope ps(HLSD_PROF_PANEL, st);
            const int pthreads = std::max(128, (n - o) / 2);
            // measured at N = 1024 (ms per 1024 matrices, all panels): one row per thread 7.75, two rows 6.70, two rows with
            // 16-column sub-panels (64 registers, two CTAs per SM) 7.11; the rank-32 update between sub-panels on the tensor
            // core (each thread's 32 multipliers = one swizzle-atom row of a 128-row operand tile, N = remaining columns,
            // product read back from TMEM and subtracted with RED, one pipeline per group of four warps): correct, 30 % fewer
            // instructions, and 7.59 - instruction fetch (`no_instruction` 3.1 warps per issue) and the serial
            // write/MMA/read-back rounds cost more than the 3000 FFMAs per row they replace
            lu_cpanel_kernel<2, 32><<<nb, pthreads, (size_t)(pthreads / 32) * kXposeWarpBytes, st>>>(PWT, M2, U, P, pm(p), pm(p + 1), n, o, pivoting);
            count_launch();
        }
        if (ct == 0) break;
        {
            ProfScope ps(HLSD_PROF_DIAG, st);
            cudaError_t e = lu_inverse_l11(M2, pm(p + 1), W, n, o, batch, st);
            if (e != cudaSuccess) return e;
        }
        // Measured alternatives of the long-K kernels (ms per 1024 matrices, both forms, all panels; default 5.24 + 1.27 for
        // the solve): the solve U12 = W X12 fused into this kernel's epilogue (X12 written to shared memory as the N operand
        // of a second product, K in two halves) 6.85 - with one CTA per SM nothing overlaps the serial epilogue; 256
        // threads, one shared-memory stage, two CTAs per SM 6.45 - the overlap of a CTA's own stores with its own MMAs is
        // worth more than a second CTA; two instead of three chunks of loads in flight 5.32; two 128 x 128 accumulators
        // instead of one 128 x 256 5.39.
        {
            ProfScope ps(HLSD_PROF_UPDATE, st);
            lu_ll_kernel<LL_ROW><<<dim3((ct + 1) / 2, nb), 512, LlSmem::kTotal, st>>>(A, M2, UT, pm(p + 1), T0, n, o);
            count_launch();
        }
        {
            ProfScope ps(HLSD_PROF_TRSM, st);
            lu_trsm_kernel<<<dim3(ct, nb), 256, MmaSmem::kTotal, st>>>(T0, W, UT, U, n, o);
            count_launch();
        }
    }
    {
        ProfScope ps(HLSD_PROF_PERMUTE, st);
        lu_assemble_l_kernel<<<dim3(panels, nb), 256, 0, st>>>(M2, L, P, hist, n, panels);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t lu_blocked(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, cudaStream_t st) {
    if (n % NB != 0 || n < 2 * NB || n > 1024) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L | (uintptr_t)U) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = ensure_dynamic_smem(lu_ll_kernel<LL_COL>, LlSmem::kTotal)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(lu_ll_kernel<LL_ROW>, LlSmem::kTotal)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(lu_trsm_kernel, MmaSmem::kTotal)) != cudaSuccess) return e;
    if ((e = ensure_dynamic_smem(lu_cpanel_kernel<2, 32>, 16 * kXposeWarpBytes + 16 * 1024)) != cudaSuccess) return e;  // (+ static)
    const long nn = (long)n * n;
    const int panels = n / NB;
    // workspace per matrix: M2, UT (n x n), PWT, T0 (n x 128), W (128 x 128), posmap history; the batch is walked in
    // slices so that the workspace stays bounded (2048 matrices of 1024 x 1024: 18 GB)
    const size_t per_mat = sizeof(float) * (size_t)(2 * nn + 2 * (long)n * NB + NB * NB) + sizeof(int) * (size_t)(panels + 1) * n;
    const long slice = std::max<long>(1, std::min<long>(std::min<long>(batch, 32768), (long)(((size_t)18 << 30) / per_mat)));
    float* ws = nullptr;
    if ((e = ws_alloc(reinterpret_cast<void**>(&ws), per_mat * (size_t)slice, st)) != cudaSuccess) return e;
    for (long b0 = 0; b0 < batch && e == cudaSuccess; b0 += slice) {
        const long cnt = std::min(slice, batch - b0);
        e = lu_crout_slice(A + b0 * nn, L + b0 * nn, U + b0 * nn, P + b0 * n, n, cnt, pivoting, ws, st);
    }
    ws_free(ws, st);
    return e;
}

}  // namespace hlsd
