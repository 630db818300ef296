// Blocked right-looking Cholesky for large N (N a multiple of 128), batched.
//
// The reference's systolic Cholesky is a right-looking elimination (Cholesky/cholesky_v1.3/
// model4x4/chol.cpp:25-63): per column k a square root, a column scaling and a rank-1 update of
// the trailing matrix.  At N = 1024 the rank-1 updates are >99 % of the work and, grouped over a
// panel of NB = 128 columns, they are a dense contraction  C -= P * P^T  (P = the finished panel
// below the diagonal block).  That part runs on the tcgen05 tensor cores; everything that is not a
// dense contraction stays on CUDA cores:
//
//   per panel p (columns o = 128 p .. o+127), all matrices of the batch at once (panel 0 reads its
//   inputs straight from A, so the operand is never copied):
//     1. blk_diag_kernel   (SIMT, register-tiled)          L11 = chol(A11), W = inv(L11)
//     2. blk_mma_kernel<TRSM>  (tcgen05)  L21 = A21 * W^T        (rows below the diagonal block)
//     3. blk_mma_kernel<SYRK>  (tcgen05)  A22 -= L21 * L21^T     (lower block triangle only)
//
// Tensor-core arithmetic: fp32 operands are split on the fly into tf32 hi (truncation) + tf32 lo parts and each
// product is formed as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (3xTF32) with fp32 accumulation in TMEM,
// which keeps ~21 mantissa bits (a single tf32 pass keeps 10).  The result is not bit-identical to
// the csim (different summation order, split products) but agrees to fp32 round-off level; the
// exact paths (HLSD_PATH_GLOBAL) remain available at any N.
//
// blk_mma_kernel: one CTA = one 128x128 output tile, D = A * B^T with A = 128 rows x 128 k,
// B = 128 rows x 128 k, both K-major fp32 in global memory.  K is consumed in four chunks of 32:
// cp.async (LDGSTS) copies the raw fp32 chunk of A and B straight into shared memory in the
// canonical K-major SWIZZLE_128B layout, double-buffered so that chunk c+1 streams in while chunk
// c is being multiplied; the raw tile itself is the tf32 "hi" operand (the tensor core reads the
// upper 19 bits of each word) and the threads only add the "lo" tile (x - hi, rounded to tf32);
// 96 KB per CTA, two CTAs per SM overlap each other's MMA and epilogue phases; one elected
// thread issues 12 tcgen05.mma.kind::tf32 (M128 N128 K8) per chunk; completion is tracked with
// tcgen05.commit on an mbarrier; the epilogue reads the accumulator with tcgen05.ld (32x32b.x32)
// and either stores it (TRSM) or subtracts it from C in place (SYRK).
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

constexpr int NB = 128;  // panel width = MMA tile edge

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M = 128, N = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

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
// One CTA of 512 threads per matrix; warp `tr` (16 warps) owns rows i = tr + 16 a (a < 8), lane
// `tc` owns columns j = tc + 32 b (b < 4): an 8 x 4 register tile, dealt cyclically so that the
// shrinking active region stays balanced.  Column j of the tile holds A[:, j] while k < j and, once
// column j has been finished (k >= j), W[:, j]: the two rank-1 updates of step k
//      A[i][j] -= l_ik * l_jk   (i > k, j > k)          W[i][j] -= l_ik * w_kj   (i > k, j <= k)
// touch disjoint columns, so one register tile and one broadcast vector u (u[j] = l_jk for j > k,
// w_kj for j <= k) serve both.  W follows from forward substitution on the identity:
// W[k][:] = (e_k - sum_{j<k} l_kj W[j][:]) / l_kk.  Per step: the owners of column k scale it,
// publish it (the systolic "column stream"), store it to L and re-seed the registers with e_k; the
// owners of row k of W scale and publish it; everyone applies the update; the owner of (k+1, k+1)
// takes the next square root.  Plain fp32 FMA arithmetic (tensor path: tolerance-checked).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2)
blk_diag_kernel(const float* __restrict__ S, float* __restrict__ L, float* __restrict__ W, int n, int o) {
    __shared__ float ubuf[NB];
    __shared__ float sdiag;
    const int tr = threadIdx.x >> 5, tc = threadIdx.x & 31;
    float* blk = L + (long)blockIdx.x * n * n + (long)o * n + o;
    const float* sblk = S + (long)blockIdx.x * n * n + (long)o * n + o;  // S = A for panel 0, L afterwards
    float v[8][4];
#pragma unroll
    for (int ia = 0; ia < 8; ia++)
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
            const int i = tr + 16 * ia, j = tc + 32 * ib;
            v[ia][ib] = (j <= i) ? sblk[(long)i * n + j] : 0.0f;
        }
    if (threadIdx.x == 0) sdiag = sqrtf(v[0][0]);
    __syncthreads();
    // k = 16 * KA + kr with KA compile-time: the register that holds row/column k is then a constant
    // index (a runtime index would push the tile into local memory)
    static_for<0, 8>([&](auto kac) {
        constexpr int KA = decltype(kac)::value;  // local row of row k
        constexpr int KB = KA >> 1;               // local column of column k
        for (int kr = 0; kr < 16; kr++) {
            const int k = 16 * KA + kr;
            const int kc = k & 31;
            const float d = sdiag;
            const float rd = 1.0f / d;
            if (tc == kc) {  // column k: scale, publish, store, re-seed with column k of the identity
#pragma unroll
                for (int ia = 0; ia < 8; ia++) {
                    const int i = tr + 16 * ia;
                    if (i > k) {
                        const float l = v[ia][KB] * rd;
                        ubuf[i] = l;
                        blk[(long)i * n + k] = l;
                        v[ia][KB] = 0.0f;
                    } else if (i == k) {
                        blk[(long)i * n + k] = d;
                        v[ia][KB] = 1.0f;
                    } else {
                        blk[(long)i * n + k] = 0.0f;  // above the diagonal (SYRK left scratch there)
                        v[ia][KB] = 0.0f;
                    }
                }
            }
            if (tr == kr) {  // row k of W (columns <= k): scale and publish
#pragma unroll
                for (int ib = 0; ib < 4; ib++) {
                    const int j = tc + 32 * ib;
                    if (j <= k) {
                        v[KA][ib] *= rd;
                        ubuf[j] = v[KA][ib];
                    }
                }
            }
            __syncthreads();
            float u[4];
#pragma unroll
            for (int ib = 0; ib < 4; ib++) u[ib] = ubuf[tc + 32 * ib];
#pragma unroll
            for (int ia = KA; ia < 8; ia++) {  // rows i = tr + 16 ia <= k for ia < KA: nothing to do
                const int i = tr + 16 * ia;
                if (i > k) {  // warp-uniform
                    const float lr = ubuf[i];
#pragma unroll
                    for (int ib = 0; ib < 4; ib++) v[ia][ib] = fmaf(-lr, u[ib], v[ia][ib]);
                }
            }
            // the owner of (k+1, k+1) holds its final value now: take the next square root
            if (k + 1 < NB && tr == ((k + 1) & 15) && tc == ((k + 1) & 31)) {
                constexpr int KA1 = KA + 1 < 8 ? KA + 1 : 7;
                sdiag = sqrtf(kr < 15 ? v[KA][KB] : v[KA1][KA1 >> 1]);
            }
            __syncthreads();
        }
    });
    float* wout = W + (long)blockIdx.x * NB * NB;
#pragma unroll
    for (int ia = 0; ia < 8; ia++)
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
            const int i = tr + 16 * ia, j = tc + 32 * ib;
            wout[i * NB + j] = (j <= i) ? v[ia][ib] : 0.0f;
        }
}

// ------------------------------------------------------------------------------------------------
// tensor-core tile kernel
// ------------------------------------------------------------------------------------------------
enum { MMA_TRSM = 0, MMA_SYRK = 1 };

struct MmaSmem {
    // per 32-wide K chunk an operand tile is 128 rows x 32 k fp32 = one swizzle atom [128 rows][128 B].
    // raw fp32 tiles (2 stages x {A, B}) are filled by cp.async and double as the tf32 "hi" operands:
    // the tensor core reads only the upper 19 bits of each 32-bit container, i.e. hi = trunc_tf32(x);
    // the "lo" tiles hold rna_tf32(x - hi) computed by the threads.
    static constexpr int kChunkK = 32;
    static constexpr int kTileBytes = 128 * kChunkK * 4;
    static constexpr int kStages = 1;  // raw stages: 1 -> 64 KB per CTA (3 CTAs/SM), 2 -> 96 KB (2 CTAs/SM)
    static constexpr int kRawA = 0, kRawB = kStages * kTileBytes;      // [stage] each
    static constexpr int kLoA = 2 * kStages * kTileBytes, kLoB = kLoA + kTileBytes;
    static constexpr int kBar = kLoB + kTileBytes;
    static constexpr int kTmemPtr = kBar + 8;
    static constexpr int kTotal = kBar + 16 + 1024;  // + slack for the 1024-byte alignment of the tiles
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// lo part of a float4 sitting in a raw tile: x - trunc_tf32(x), rounded to tf32
__device__ __forceinline__ float4 lo_of(float4 v) {
    float4 l;
    l.x = to_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
    l.y = to_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
    l.z = to_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
    l.w = to_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
    return l;
}

// grid: (tiles, batch).  TRSM: tile t = row block t of the panel below the diagonal block.
//                         SYRK: tile t -> (tj >= ti) in the lower block triangle of the trailing matrix.
template <int MODE>
__global__ void __launch_bounds__(256, MmaSmem::kStages == 1 ? 3 : 2)
blk_mma_kernel(const float* __restrict__ S, float* __restrict__ L, const float* __restrict__ W, int n, int o) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MmaSmem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MmaSmem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float* Lb = L + (long)blockIdx.y * n * n;
    const float* Sb = S + (long)blockIdx.y * n * n;  // S = A for panel 0 (inputs not yet in L), L afterwards

    int tj, ti;
    if constexpr (MODE == MMA_TRSM) {
        tj = blockIdx.x;
        ti = 0;
    } else {
        // unrank blockIdx.x into (tj, ti), ti <= tj
        int r = blockIdx.x;
        tj = (int)((sqrtf(8.0f * r + 1.0f) - 1.0f) * 0.5f);
        while ((tj + 1) * (tj + 2) / 2 <= r) tj++;
        while (tj * (tj + 1) / 2 > r) tj--;
        ti = r - tj * (tj + 1) / 2;
    }
    const int row0 = o + NB;  // first row of the panel below the diagonal block / of the trailing matrix
    // operand rows (K-major, k = panel columns o .. o+127)
    // TRSM multiplies the not yet factored panel rows (from S); SYRK multiplies finished L21 rows (from L)
    const float* Arows = (MODE == MMA_TRSM ? Sb : Lb) + (long)(row0 + tj * NB) * n + o;
    const long lda = n;
    const float* Brows;
    long ldb;
    if constexpr (MODE == MMA_TRSM) {
        Brows = W + (long)blockIdx.y * NB * NB;  // B[i][k] = W[i][k]  ->  D = A21 * W^T
        ldb = NB;
    } else {
        Brows = Lb + (long)(row0 + ti * NB) * n + o;
        ldb = n;
    }
    const bool same = (MODE == MMA_SYRK) && (tj == ti);

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
            if (!same) cp_async16(rb + piece_off(it), Brows + row * ldb + chunk * MmaSmem::kChunkK + c4 * 4);
        }
        cp_async_commit();
    };
    issue_loads(0);
    for (int chunk = 0; chunk < kChunks; chunk++) {
        unsigned char* ra = smem + MmaSmem::kRawA + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
        unsigned char* rb = same ? ra : smem + MmaSmem::kRawB + (chunk % MmaSmem::kStages) * MmaSmem::kTileBytes;
        unsigned char* la = smem + MmaSmem::kLoA;
        unsigned char* lb = same ? la : smem + MmaSmem::kLoB;
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
            if (!same) *reinterpret_cast<float4*>(lb + off) = lo_of(*reinterpret_cast<const float4*>(rb + off));
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
    // coalesced read-modify-write of C, one 512-byte row per warp instruction.
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
        float* Cbase;
        if constexpr (MODE == MMA_TRSM) Cbase = Lb + (long)(row0 + tj * NB) * n + o;       // L21
        else Cbase = Lb + (long)(row0 + tj * NB) * n + (row0 + ti * NB);                   // C tile
        const float* Csrc = Sb + (long)(row0 + tj * NB) * n + (row0 + ti * NB);            // SYRK reads C from S
#pragma unroll 4
        for (int rr = 0; rr < 16; rr++) {
            const int row = warp * 16 + rr;
            const float4 d = *reinterpret_cast<const float4*>(stage + row * 128 + ((lane ^ (row & 7)) << 2));
            float4* cp = reinterpret_cast<float4*>(Cbase + (long)row * n) + lane;
            float4 v;
            if constexpr (MODE == MMA_TRSM) {
                v = d;
            } else {
                v = *(reinterpret_cast<const float4*>(Csrc + (long)row * n) + lane);
                v.x -= d.x, v.y -= d.y, v.z -= d.z, v.w -= d.w;
            }
            *cp = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 128);
}

// ------------------------------------------------------------------------------------------------
cudaError_t cholesky_blocked(const float* A, float* L, int n, long batch, int use_tensor, cudaStream_t st) {
    if (n % NB != 0 || n < 2 * NB || !use_tensor) return cudaErrorNotSupported;
    if ((((uintptr_t)A | (uintptr_t)L) & 15) != 0) return cudaErrorNotSupported;
    if (batch == 0) return cudaSuccess;
    if (batch > 65535) return cudaErrorNotSupported;
    cudaError_t e;
    float* W = nullptr;
    if ((e = cudaMallocAsync(&W, sizeof(float) * (size_t)batch * NB * NB, st)) != cudaSuccess) return e;
    cudaFuncSetAttribute(blk_mma_kernel<MMA_TRSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, MmaSmem::kTotal);
    cudaFuncSetAttribute(blk_mma_kernel<MMA_SYRK>, cudaFuncAttributeMaxDynamicSharedMemorySize, MmaSmem::kTotal);
    const int panels = n / NB;
    blk_zero_upper_kernel<<<dim3(panels * (panels - 1) / 2, (unsigned)batch), 256, 0, st>>>(L, n);
    count_launch();
    for (int p = 0; p < panels; p++) {
        const int o = p * NB;
        const int mt = panels - p - 1;  // 128-row blocks below the diagonal block
        const float* src = (p == 0) ? A : L;
        blk_diag_kernel<<<(unsigned)batch, 512, 0, st>>>(src, L, W, n, o);
        count_launch();
        if (mt > 0) {
            blk_mma_kernel<MMA_TRSM><<<dim3(mt, (unsigned)batch), 256, MmaSmem::kTotal, st>>>(src, L, W, n, o);
            count_launch();
            blk_mma_kernel<MMA_SYRK><<<dim3(mt * (mt + 1) / 2, (unsigned)batch), 256, MmaSmem::kTotal, st>>>(src, L, W, n, o);
            count_launch();
        }
    }
    e = cudaGetLastError();
    cudaFreeAsync(W, st);
    return e;
}

}  // namespace hlsd
