// Blocked right-looking Cholesky for large N (N a multiple of 128), batched.
//
// The reference's systolic Cholesky is a right-looking elimination (Cholesky/cholesky_v1.3/
// model4x4/chol.cpp:25-63): per column k a square root, a column scaling and a rank-1 update of
// the trailing matrix.  At N = 1024 the rank-1 updates are >99 % of the work and, grouped over a
// panel of NB = 128 columns, they are a dense contraction  C -= P * P^T  (P = the finished panel
// below the diagonal block).  That part runs on the tcgen05 tensor cores; everything that is not a
// dense contraction stays on CUDA cores:
//
//   per panel p (columns o = 128 p .. o+127), all matrices of the batch at once:
//     1. blk_diag_kernel   (SIMT, register-tiled)          L11 = chol(A11), W = inv(L11)
//     2. blk_mma_kernel<TRSM>  (tcgen05)  L21 = A21 * W^T        (rows below the diagonal block)
//     3. blk_mma_kernel<SYRK>  (tcgen05)  A22 -= L21 * L21^T     (lower block triangle only)
//
// Tensor-core arithmetic: fp32 operands are split on the fly into tf32 hi + tf32 lo parts and each
// product is formed as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (3xTF32) with fp32 accumulation in TMEM,
// which keeps ~21 mantissa bits (a single tf32 pass keeps 10).  The result is not bit-identical to
// the csim (different summation order, split products) but agrees to fp32 round-off level; the
// exact paths (HLSD_PATH_GLOBAL) remain available at any N.
//
// blk_mma_kernel: one CTA = one 128x128 output tile, D = A * B^T with A = 128 rows x 128 k,
// B = 128 rows x 128 k, both K-major fp32 in global memory.  K is consumed in four chunks of 32:
// 256 threads load the chunk with coalesced 16-byte loads, split hi/lo in registers and store the
// four operand tiles (64 KB together, so three CTAs share an SM and overlap each other's load, MMA
// and epilogue phases) into shared memory in the canonical K-major SWIZZLE_128B layout (the layout
// TMA would produce; TMA itself cannot split, so the staging is done by the threads); one elected
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
// L = tril(A): the factorisation runs in place in L
// ------------------------------------------------------------------------------------------------
__global__ void blk_init_kernel(const float* __restrict__ A, float* __restrict__ L, int n, long total4) {
    const long stride = (long)gridDim.x * blockDim.x;
    const int n4 = n / 4;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += stride) {
        const long rowid = e / n4;
        const int c4 = (int)(e - rowid * n4);
        const int i = (int)(rowid % n);
        float4 v = reinterpret_cast<const float4*>(A)[e];
        const int j = 4 * c4;
        if (j + 0 > i) v.x = 0.0f;
        if (j + 1 > i) v.y = 0.0f;
        if (j + 2 > i) v.z = 0.0f;
        if (j + 3 > i) v.w = 0.0f;
        reinterpret_cast<float4*>(L)[e] = v;
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
__global__ void __launch_bounds__(512, 1)
blk_diag_kernel(float* __restrict__ L, float* __restrict__ W, int n, int o) {
    __shared__ float ubuf[NB];
    __shared__ float sdiag;
    const int tr = threadIdx.x >> 5, tc = threadIdx.x & 31;
    float* blk = L + (long)blockIdx.x * n * n + (long)o * n + o;
    float v[8][4];
#pragma unroll
    for (int ia = 0; ia < 8; ia++)
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
            const int i = tr + 16 * ia, j = tc + 32 * ib;
            v[ia][ib] = (j <= i) ? blk[(long)i * n + j] : 0.0f;
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
    // four operand tiles, each 128 rows x 32 k fp32 = one swizzle atom of [128 rows][128 B]
    static constexpr int kChunkK = 32;
    static constexpr int kTileBytes = 128 * kChunkK * 4;
    static constexpr int kAhi = 0, kAlo = kTileBytes, kBhi = 2 * kTileBytes, kBlo = 3 * kTileBytes;
    static constexpr int kBar = 4 * kTileBytes;
    static constexpr int kTmemPtr = kBar + 8;
    static constexpr int kTotal = kBar + 16 + 1024;  // + slack for the 1024-byte alignment of the tiles
};

__device__ __forceinline__ void split_store(unsigned char* hi_tile, unsigned char* lo_tile, int row, int c4, float4 v) {
    // 16-byte chunk c4 (4 k) of `row` inside a 128 x 32 tile; SWIZZLE_128B: chunk index ^= row % 8
    const int off = row * 128 + ((c4 ^ (row & 7)) << 4);
    float4 h, l;
    h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
    l.x = to_tf32(v.x - h.x), l.y = to_tf32(v.y - h.y), l.z = to_tf32(v.z - h.z), l.w = to_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// grid: (tiles, batch).  TRSM: tile t = row block t of the panel below the diagonal block.
//                         SYRK: tile t -> (tj >= ti) in the lower block triangle of the trailing matrix.
template <int MODE>
__global__ void __launch_bounds__(256, 2)
blk_mma_kernel(float* __restrict__ L, const float* __restrict__ W, int n, int o) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MmaSmem::kBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MmaSmem::kTmemPtr);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float* Lb = L + (long)blockIdx.y * n * n;

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
    const float* Arows = Lb + (long)(row0 + tj * NB) * n + o;
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

    unsigned char* sAhi = smem + MmaSmem::kAhi;
    unsigned char* sAlo = smem + MmaSmem::kAlo;
    unsigned char* sBhi = same ? sAhi : smem + MmaSmem::kBhi;
    unsigned char* sBlo = same ? sAlo : smem + MmaSmem::kBlo;

    constexpr int kChunks = NB / MmaSmem::kChunkK;  // 4 chunks of 32 k
    float4 ra[4], rb[4];
    auto gload = [&](int chunk) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
            ra[it] = *reinterpret_cast<const float4*>(Arows + row * lda + chunk * MmaSmem::kChunkK + c4 * 4);
            if (!same) rb[it] = *reinterpret_cast<const float4*>(Brows + row * ldb + chunk * MmaSmem::kChunkK + c4 * 4);
        }
    };
    gload(0);
    for (int chunk = 0; chunk < kChunks; chunk++) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int idx = it * 256 + t, row = idx >> 3, c4 = idx & 7;
            split_store(sAhi, sAlo, row, c4, ra[it]);
            if (!same) split_store(sBhi, sBlo, row, c4, rb[it]);
        }
        fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (t == 0) {
            const uint64_t dAhi = make_sw128_desc(smem_u32(sAhi)), dAlo = make_sw128_desc(smem_u32(sAlo));
            const uint64_t dBhi = make_sw128_desc(smem_u32(sBhi)), dBlo = make_sw128_desc(smem_u32(sBlo));
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
        if (chunk + 1 < kChunks) gload(chunk + 1);  // next chunk's global loads fly while the tensor core works
        mbar_wait(bar, (uint32_t)(chunk & 1));  // MMAs done: operand tiles may be overwritten / accumulator is final
        tc_fence_after();
    }

    // epilogue: thread -> accumulator row 32*(warp%4)+lane, 64 columns (warp/4)
    const int drow = (warp & 3) * 32 + lane;
    const int dcol0 = (warp >> 2) * 64;
    float* Crow;
    if constexpr (MODE == MMA_TRSM) Crow = Lb + (long)(row0 + tj * NB + drow) * n + o;                      // overwrite A21
    else Crow = Lb + (long)(row0 + tj * NB + drow) * n + (row0 + ti * NB);                                  // C tile
#pragma unroll
    for (int h = 0; h < 2; h++) {
        float acc[32];
        const int col = dcol0 + h * 32;
        tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col, acc);
        float4* c4p = reinterpret_cast<float4*>(Crow + col);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float4 v;
            if constexpr (MODE == MMA_TRSM) {
                v = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            } else {
                v = c4p[q];
                v.x -= acc[4 * q], v.y -= acc[4 * q + 1], v.z -= acc[4 * q + 2], v.w -= acc[4 * q + 3];
            }
            c4p[q] = v;
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
    const long total4 = batch * (long)n * n / 4;
    blk_init_kernel<<<(unsigned)std::min<long>((total4 + 255) / 256, (long)sm_count() * 16), 256, 0, st>>>(A, L, n, total4);
    count_launch();
    const int panels = n / NB;
    for (int p = 0; p < panels; p++) {
        const int o = p * NB;
        const int mt = panels - p - 1;  // 128-row blocks below the diagonal block
        blk_diag_kernel<<<(unsigned)batch, 512, 0, st>>>(L, W, n, o);
        count_launch();
        if (mt > 0) {
            blk_mma_kernel<MMA_TRSM><<<dim3(mt, (unsigned)batch), 256, MmaSmem::kTotal, st>>>(L, W, n, o);
            count_launch();
            blk_mma_kernel<MMA_SYRK><<<dim3(mt * (mt + 1) / 2, (unsigned)batch), 256, MmaSmem::kTotal, st>>>(L, W, n, o);
            count_launch();
        }
    }
    e = cudaGetLastError();
    cudaFreeAsync(W, st);
    return e;
}

}  // namespace hlsd
