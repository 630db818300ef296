// Batched LU, the reference's `int top(A, L, U, P)` (LU/lu_v*/model4x4/lu.h:19-22): Doolittle
// elimination with partial pivoting, LINPACK-style row exchanges (rows k and P[k] are exchanged in
// columns >= k only; L columns already produced are not touched), lu_v1.1/model4x4/lu.cpp:23-92.
// lu_v1.2 and lu_v2.1 reorder the same per-element operations (checked bit for bit in the CPU
// tests), so one arithmetic serves the three designs.  `pivoting == 0` skips the search (P[k] = k).
//
// Kernels:
//   lu_tpm_kernel<N>      thread-per-matrix in registers, N = 4: staged slabs.
//   lu_regs_kernel<N,R>   R lanes per matrix, all in registers, N in {8, 16}: pivot row by shuffle.
//   lu_lane_kernel<N>     one lane per row, N = 32 (one warp per matrix), 64 (two) and 128 (four).
//   lu_group_kernel<G>    G threads per matrix, matrix resident in shared memory.
//   lu_global_kernel      one CTA per matrix, in place in global memory (any N; exact; slow).
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

// Pivot candidate ordering = the reference's downward scan with a strict '<' (lu.cpp:44-52):
// the smallest row index among the largest |a|.  NaNs below the diagonal never win; a NaN on
// the diagonal is never displaced (every comparison with it is false).
struct PivotCand {
    float v;
    int i;
};
__device__ __forceinline__ PivotCand pivot_better(PivotCand a, PivotCand b) {
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ float pivot_key(float x, bool is_diag) {
    float v = fabsf(x);
    if (v != v) v = is_diag ? __int_as_float(0x7f800000) : -1.0f;
    return v;
}

// ------------------------------------------------------------------------------------------------
// thread-per-matrix (registers), slabs staged by bulk copies as in chol_tpm_kernel
// ------------------------------------------------------------------------------------------------
template <int N, int WARPS, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, 1)
lu_tpm_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ U, int* __restrict__ P,
              long batch, int pivoting) {
    constexpr int kSlot = N * N + 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // per warp: slab for A (reused for L) and a slab for U
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * 2 * 32 * kSlot * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* slotA = slabs + ((size_t)warp * 2 * 32 + lane) * kSlot;
    float* slotU = slotA + 32 * kSlot;
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
        constexpr uint32_t kMatBytes = N * N * 4;
        float4* slabA = reinterpret_cast<float4*>(slabs + (size_t)warp * 2 * 32 * kSlot);
        float4* slabU = slabA + 32 * (kSlot / 4);
        if constexpr (kCoalesced) {
            __syncwarp();
            slab_g2s<N * N / 4>(slabA, reinterpret_cast<const float4*>(A + m0 * (long)(N * N)), valid, lane);
            __syncwarp();
        } else {
            bulk_wait_read<0>();
            if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * kMatBytes);
            __syncwarp();
            if (lane < valid) bulk_g2s(slotA, A + (m0 + lane) * (long)(N * N), kMatBytes, bar);
            mbar_wait(bar, parity);
            parity ^= 1;
        }
        if (lane < valid) {
            float a[N][N];
            const float4* s4 = reinterpret_cast<const float4*>(slotA);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int q = 0; q < N / 4; q++) {
                    float4 v = s4[i * (N / 4) + q];
                    a[i][4 * q] = v.x, a[i][4 * q + 1] = v.y, a[i][4 * q + 2] = v.z, a[i][4 * q + 3] = v.w;
                }
            int piv[N];
#pragma unroll
            for (int k = 0; k < N - 1; k++) {
                int p = k;
                if (pivoting) {
                    float maxpwr = fabsf(a[k][k]);
#pragma unroll
                    for (int i = k + 1; i < N; i++) {
                        float v = fabsf(a[i][k]);
                        if (maxpwr < v) {
                            maxpwr = v;
                            p = i;
                        }
                    }
                }
                piv[k] = p;
                // exchange rows k and p in columns >= k (register arrays: select, no dynamic index)
#pragma unroll
                for (int i = k + 1; i < N; i++)
                    if (p == i) {
#pragma unroll
                        for (int j = k; j < N; j++) {
                            float t = a[i][j];
                            a[i][j] = a[k][j];
                            a[k][j] = t;
                        }
                    }
                float diag = a[k][k];
#pragma unroll
                for (int i = k + 1; i < N; i++) a[i][k] = xdiv(a[i][k], diag);
#pragma unroll
                for (int i = k + 1; i < N; i++)
#pragma unroll
                    for (int j = k + 1; j < N; j++) a[i][j] = msub<FUSED>(a[i][j], a[i][k], a[k][j]);
            }
            piv[N - 1] = N - 1;
            float4* l4 = reinterpret_cast<float4*>(slotA);
            float4* u4 = reinterpret_cast<float4*>(slotU);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int q = 0; q < N / 4; q++) {
                    float l[4], u[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        int j = 4 * q + t;
                        l[t] = (j < i) ? a[i][j] : (j == i ? 1.0f : 0.0f);
                        u[t] = (j >= i) ? a[i][j] : 0.0f;
                    }
                    l4[i * (N / 4) + q] = make_float4(l[0], l[1], l[2], l[3]);
                    u4[i * (N / 4) + q] = make_float4(u[0], u[1], u[2], u[3]);
                }
            int4* p4 = reinterpret_cast<int4*>(P + (m0 + lane) * (long)N);
#pragma unroll
            for (int q = 0; q < N / 4; q++) p4[q] = make_int4(piv[4 * q], piv[4 * q + 1], piv[4 * q + 2], piv[4 * q + 3]);
            if constexpr (!kCoalesced) {
                fence_async_smem();
                bulk_s2g(L + (m0 + lane) * (long)(N * N), slotA, kMatBytes);
                bulk_s2g(U + (m0 + lane) * (long)(N * N), slotU, kMatBytes);
                bulk_commit();
            }
        }
        if constexpr (kCoalesced) {
            __syncwarp();
            slab_s2g<N * N / 4>(reinterpret_cast<float4*>(L + m0 * (long)(N * N)), slabA, valid, lane);
            slab_s2g<N * N / 4>(reinterpret_cast<float4*>(U + m0 * (long)(N * N)), slabU, valid, lane);
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// R lanes per matrix, everything in registers (N = 8, 16, 32): lane r of a group owns data rows
// r + R t, each a full row of N registers; the step loop is unrolled at compile time, so registers are
// only ever indexed by constants.  Rows are never moved: every data row carries its current position
// pos[t]; a pivot exchange swaps two positions.  Per step k:
//   pivot search   each lane offers (|a_ik|, position) of its rows still below the pivot; butterfly
//                  arg-max over the group with the reference's tie rule (smallest position wins)
//   pivot row      its owner lane (found by a ballot over the group) selects it from its T rows and
//                  broadcasts it column by column by shuffle (the systolic "pivot broadcast"); the
//                  owner also stores it as row k of U
//   column / update  every lane divides its entries of column k by the pivot, stores them into the
//                  L slab at the rows' current positions, and subtracts l_ik * u_kj from its rows
// L (unit lower) is assembled in a shared-memory slab and leaves by a bulk store; so does U unless
// U_DIRECT.  Operation order per element is the reference's (k ascending, separate roundings):
// bit-identical.  Rows that have already served as pivot rows are dead data and are updated along with
// the live ones (no predication); only their L stores are masked.
// ------------------------------------------------------------------------------------------------
template <int N, int R, int WARPS, bool PIVOT, bool U_DIRECT, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, 1)
lu_regs_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ U, int* __restrict__ P,
               long batch) {
    constexpr int MPW = 32 / R;  // matrices per warp
    constexpr int T = N / R;     // data rows per lane
    constexpr int kSlot = N * N + 4;
    constexpr int kPerMat = (U_DIRECT ? 1 : 2) * kSlot + N;  // floats: A/L slab, U slab (unless U rows go straight out), P
    constexpr uint32_t kMatBytes = N * N * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * MPW * kPerMat * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane % R, m = lane / R, gbase = lane - r;
    float* sA = base + ((size_t)warp * MPW + m) * kPerMat;  // operand, then L
    float* sU = sA + kSlot;
    int* sP = reinterpret_cast<int*>(sA + (U_DIRECT ? 1 : 2) * kSlot);
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long num_slabs = (batch + MPW - 1) / MPW;
    uint32_t parity = 0;
    for (long s = (long)blockIdx.x * WARPS + warp; s < num_slabs; s += (long)gridDim.x * WARPS) {
        const long m0 = s * MPW;
        const int valid = (int)min((long)MPW, batch - m0);
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * kMatBytes);
        __syncwarp();
        if (m < valid && r == 0) bulk_g2s(sA, A + (m0 + m) * (long)(N * N), kMatBytes, bar);  // one copy per matrix
        mbar_wait(bar, parity);
        parity ^= 1;
        float a[T][N];
        int pos[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            pos[t] = t * R + r;
            const float4* row4 = reinterpret_cast<const float4*>(sA + (t * R + r) * N);
#pragma unroll
            for (int q = 0; q < N / 4; q++) {
                float4 v = row4[q];
                a[t][4 * q] = v.x, a[t][4 * q + 1] = v.y, a[t][4 * q + 2] = v.z, a[t][4 * q + 3] = v.w;
            }
        }
        __syncwarp();  // the operand is in registers: the slab becomes L (diagonal and above), the other U (below)
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int i = t * R + r;
            float4* l4 = reinterpret_cast<float4*>(sA + i * N);
            float4* u4 = reinterpret_cast<float4*>(sU + i * N);
#pragma unroll
            for (int q = 0; q < N / 4; q++) {
                if (q >= (t * R) / 4 && q >= i / 4)
                    l4[q] = make_float4(i == 4 * q ? 1.0f : 0.0f, i == 4 * q + 1 ? 1.0f : 0.0f, i == 4 * q + 2 ? 1.0f : 0.0f,
                                        i == 4 * q + 3 ? 1.0f : 0.0f);
                if (!U_DIRECT && q <= (t * R + R - 1) / 4 && q < i / 4) u4[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        __syncwarp();
        float u[N];
        // pivot search over column k (rows at positions >= k): butterfly arg-max over the group
        auto search = [&](auto kc) -> int {
            constexpr int k = decltype(kc)::value;
            if constexpr (!PIVOT || k >= N - 1) return k;
            PivotCand c{-2.0f, 0x7fffffff};
#pragma unroll
            for (int t = 0; t < T; t++)
                if (pos[t] >= k) c = pivot_better(c, PivotCand{pivot_key(a[t][k], pos[t] == k), pos[t]});
            // (REDUX on the group's lane mask instead of this butterfly was measured: 1.03 G -> 0.59 G fact/s at N = 16 -
            // the eight 4-lane groups of a warp execute the collective one after the other)
#pragma unroll
            for (int off = R / 2; off > 0; off >>= 1) {
                PivotCand o{__shfl_xor_sync(0xffffffffu, c.v, off), __shfl_xor_sync(0xffffffffu, c.i, off)};
                c = pivot_better(c, o);
            }
            return c.i;
        };
        int p = search(std::integral_constant<int, 0>{});
        static_for<0, N>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            if (r == 0) sP[k] = p;
            bool isp[T];
            bool mine = false;
#pragma unroll
            for (int t = 0; t < T; t++) {
                if (pos[t] == p) pos[t] = k;
                else if (pos[t] == k) pos[t] = p;
                isp[t] = pos[t] == k;
                mine |= isp[t];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            const int owner = gbase + __ffs((bal >> gbase) & (R == 32 ? 0xffffffffu : (1u << (R & 31)) - 1u)) - 1;
            // pivot row: selected from the owner's rows and broadcast over the group (the systolic "pivot broadcast")
#pragma unroll
            for (int j = k; j < N; j++) {
                float v = a[0][j];
#pragma unroll
                for (int t = 1; t < T; t++) v = isp[t] ? a[t][j] : v;
                u[j] = __shfl_sync(0xffffffffu, v, owner);
            }
            // ... and stored as row k of U: every lane holds it now, so lane q % R stores quad q - one predicated
            // 16-byte store per lane instead of N/4 masked stores by the owner alone inside a divergent block
            {
                float4* u4 = U_DIRECT ? reinterpret_cast<float4*>(U + (m0 + m) * (long)(N * N) + k * N)
                                      : reinterpret_cast<float4*>(sU + k * N);
                static_for<(U_DIRECT ? 0 : k / 4), N / 4>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    const float4 v = make_float4(4 * q >= k ? u[4 * q >= k ? 4 * q : k] : 0.0f,
                                                 4 * q + 1 >= k ? u[4 * q + 1 >= k ? 4 * q + 1 : k] : 0.0f,
                                                 4 * q + 2 >= k ? u[4 * q + 2 >= k ? 4 * q + 2 : k] : 0.0f,
                                                 4 * q + 3 >= k ? u[4 * q + 3] : 0.0f);
                    if (r == q % R) {
                        if constexpr (U_DIRECT) {
                            if (m < valid) __stcs(u4 + q, v);
                        } else {
                            u4[q] = v;
                        }
                    }
                });
            }
            if constexpr (k < N - 1) {
                const float piv = u[k];
                const SharedDivisor dv = make_divisor(piv);
                // dead rows divide the pivot by itself: a zero numerator would take the slow division path
                float num[T], lm[T];
                bool ok = dv.ok;
#pragma unroll
                for (int t = 0; t < T; t++) {
                    num[t] = pos[t] > k ? a[t][k] : piv;
                    ok &= div_in_range(num[t]);
                }
                if (ok) {
#pragma unroll
                    for (int t = 0; t < T; t++) lm[t] = fast_quotient(num[t], dv);
                } else {
#pragma unroll
                    for (int t = 0; t < T; t++) lm[t] = xdiv(num[t], piv);
                }
                // column k+1 first: the next step's pivot search only needs that one and its shuffles
                // overlap the rest of the update (one straight-line block, the scheduler interleaves)
#pragma unroll
                for (int t = 0; t < T; t++) a[t][k + 1] = msub<FUSED>(a[t][k + 1], lm[t], u[k + 1]);
                p = search(std::integral_constant<int, k + 1>{});
#pragma unroll
                for (int t = 0; t < T; t++) {
                    if (pos[t] > k) sA[pos[t] * N + k] = lm[t];
#pragma unroll
                    for (int j = k + 2; j < N; j++) a[t][j] = msub<FUSED>(a[t][j], lm[t], u[j]);
                }
            }
        });
        fence_async_smem();
        __syncwarp();
        if (m < valid) {
            // one store per matrix and output, issued by different lanes of its group
            if (r == 0) bulk_s2g(L + (m0 + m) * (long)(N * N), sA, kMatBytes);
            if (!U_DIRECT && r == 1) bulk_s2g(U + (m0 + m) * (long)(N * N), sU, kMatBytes);
            if (r == R - 1) bulk_s2g(P + (m0 + m) * (long)N, sP, N * 4);
            bulk_commit();
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// One lane per row (N = 64: two warps per matrix, N = 128: four): position tracking as in lu_regs_kernel with a
// single data row per lane, the step loop unrolled at compile time.  ONE named barrier per column: every warp
// reduces its own candidates by butterfly, and its local winner publishes (candidate, data row) in the warp's
// slot of a buffer that is double-buffered by the parity of k; after the barrier every lane reduces the
// WPM slot candidates itself and reads the winning slot's row.  The butterfly of the next column's search is
// issued right after that column is updated and overlaps the rest of the update.
// Round-2 changes, from the ncu capture of the round-1 kernel (profiles/r02_lu64_baseline_ncu_summary.txt: 43 %
// of the shared-memory wavefronts were bank conflicts, issue slots 40 % busy with `short_scoreboard`, `barrier`
// and `mio_throttle` on top): (1) the multipliers are scattered into a slab with an ODD row stride (N + 1 floats:
// lanes at distinct positions hit distinct banks; with stride N all 32 stores of a column went to one bank);
// the slab is written out row by row by the matrix's own threads, which also supply the unit diagonal and the
// zeros, so it needs no initialisation; (2) row k of U is stored by all N threads together (one coalesced
// element each) instead of by the pivot lane alone inside a divergent block (16 masked float4 stores per
// column executed for one lane); (3) one barrier per column instead of two.
// ------------------------------------------------------------------------------------------------
// LOCKSTEP: the per-matrix barrier becomes a CTA-wide one, so that all matrices of a CTA walk through the unrolled
// code together (see chol_rows_kernel: instruction fetch is the top stall of the 128 x 128 kernel).
constexpr int kDirectLoadMaxN = 128;
template <int N, int MATS, bool PIVOT, bool FUSED, bool LOCKSTEP = false>
__global__ void __launch_bounds__(MATS * N, 1)
lu_lane_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ U, int* __restrict__ P,
               long batch) {
    static_assert(N == 32 || N == 64 || N == 128, "one, two or four warps per matrix");
    constexpr int WPM = N / 32;  // warps per matrix
    constexpr int LD = N + 1;    // odd row stride of the L slab
    constexpr int kPerMat = N * LD + 2 * WPM * N + N + 4 * WPM;  // floats: slab, row slots, P, candidates
    static_assert((N * LD) % 4 == 0 && kPerMat % 4 == 0, "16-byte alignment of the pieces");
    constexpr uint32_t kRowBytes = N * 4;
    constexpr bool kDirectLoad = N <= kDirectLoadMaxN;  // operand rows straight from global memory into registers
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)MATS * kPerMat * 4);
    const int mat = threadIdx.x / N, r = threadIdx.x % N, lane = threadIdx.x & 31, wim = r >> 5;
    float* sA = base + (size_t)mat * kPerMat;  // operand rows (dense, stride N) while loading, then the L slab (stride LD)
    float* sRow = sA + N * LD;                 // [2 parities][WPM slots][N]: candidate pivot rows
    int* sP = reinterpret_cast<int*>(sRow + 2 * WPM * N);
    PivotCand* part = reinterpret_cast<PivotCand*>(sP + N);  // [2 parities][WPM]
    uint64_t* bar = bars + mat;
    auto gsync = [&]() {
        if constexpr (LOCKSTEP) __syncthreads();
        else if constexpr (WPM == 1) __syncwarp();
        else named_bar_sync(1 + mat, N);
    };
    if (r == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    gsync();
    uint32_t parity = 0;
    for (long b0 = (long)blockIdx.x * MATS; b0 < batch; b0 += (long)gridDim.x * MATS) {
        const long b = b0 + mat;
        const bool live = b < batch;  // (LOCKSTEP: idle matrices of the last round keep step with the others)
        if (!LOCKSTEP && !live) break;
        bulk_wait_read<0>();   // the previous matrix's P has left
        float a[N];
        if constexpr (kDirectLoad) {
            // N = 32: every lane reads its own 128-byte row straight into registers.  ncu on the staged version: the
            // proxy fence in front of the bulk copies was the top stall line (10 % of the samples, `membar`), and with 32
            // warps per SM nothing is gained by staging
            gsync();
            const float4* g4 = reinterpret_cast<const float4*>(A + b * (long)(N * N) + r * N);
#pragma unroll
            for (int q = 0; q < N / 4; q++) {
                const float4 v = live ? __ldg(g4 + q) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                a[4 * q] = v.x, a[4 * q + 1] = v.y, a[4 * q + 2] = v.z, a[4 * q + 3] = v.w;
            }
        } else {
            fence_async_smem();    // ... and its slab was read with ordinary loads: order them before the bulk copies below
            gsync();
            if (r == 0) mbar_arrive_expect_tx(bar, live ? N * kRowBytes : 0);
            gsync();
            if (live) bulk_g2s(sA + r * N, A + b * (long)(N * N) + r * N, kRowBytes, bar);
            mbar_wait(bar, parity);
            parity ^= 1;
            const float4* row4 = reinterpret_cast<const float4*>(sA + r * N);
#pragma unroll
            for (int q = 0; q < N / 4; q++) {
                float4 v = row4[q];
                a[4 * q] = v.x, a[4 * q + 1] = v.y, a[4 * q + 2] = v.z, a[4 * q + 3] = v.w;
            }
        }
        float* ug = U + b * (long)(N * N);  // U rows go straight to global memory, one finished row per step
        int pos = r;
        gsync();  // operand in registers: the buffer becomes the L slab
        // this warp's best candidate of column k (rows at positions >= k)
        auto search = [&](auto kc) -> PivotCand {
            constexpr int k = decltype(kc)::value;
            PivotCand c{-2.0f, 0x7fffffff};
            if constexpr (PIVOT && k < N - 1) {
                // arg-max over the warp with two REDUX instead of a five-level butterfly: the key (-2 = no candidate,
                // -1 = NaN off the diagonal, else |a| >= 0) is mapped monotonically to an unsigned integer and maximised,
                // then the smallest position among the lanes that hold the maximum is taken (the reference's tie rule)
                unsigned uk = 0u;
                if (pos >= k) {
                    const float v = pivot_key(a[k], pos == k);
                    uk = v < 0.0f ? 1u : __float_as_uint(v) + 2u;
                }
                const unsigned wmax = __reduce_max_sync(0xffffffffu, uk);
                const unsigned wpos = __reduce_min_sync(0xffffffffu, (uk == wmax && wmax != 0u) ? (unsigned)pos : 0x7fffffffu);
                c = PivotCand{wmax >= 2u ? __uint_as_float(wmax - 2u) : (wmax == 1u ? -1.0f : -2.0f), (int)wpos};
            }
            return c;
        };
        PivotCand cand = search(std::integral_constant<int, 0>{});
        static_for<0, N>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr bool kSearch = PIVOT && k < N - 1;
            float* slots = sRow + (k & 1) * WPM * N;
            // publish: the warp's local winner (or, without a search, the lane at position k) offers its row
            const bool offer = kSearch ? (pos == cand.i) : (pos == k);
            if (offer) {
                float4* dst = reinterpret_cast<float4*>(slots + (kSearch ? wim : 0) * N);
#pragma unroll
                for (int q = k / 4; q < N / 4; q++) dst[q] = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
            }
            if constexpr (kSearch && WPM > 1) {
                if (lane == 0) part[(k & 1) * WPM + wim] = cand;
            }
            gsync();
            int p = k, slot = 0;
            if constexpr (kSearch) {
                if constexpr (WPM > 1) {
                    const PivotCand* pp = part + (k & 1) * WPM;
                    PivotCand best = pp[0];
#pragma unroll
                    for (int w = 1; w < WPM; w++) {
                        const PivotCand o = pp[w];
                        const bool take = (o.v > best.v || (o.v == best.v && o.i < best.i));
                        best = take ? o : best;
                        slot = take ? w : slot;
                    }
                    p = best.i;
                } else {
                    p = cand.i;
                }
            }
            if (r == 0) sP[k] = p;
            if (pos == p) pos = k;
            else if (pos == k) pos = p;
            const float* urow = slots + slot * N;
            if (live) __stcs(ug + k * N + r, r >= k ? urow[r] : 0.0f);  // row k of U, one element per thread, coalesced
            if constexpr (k < N - 1) {
                const float4* ur4 = reinterpret_cast<const float4*>(urow);
                const float piv = urow[k];
                // rows that have already served as pivot rows are dead data: they divide the pivot by itself
                // (a value on the fast division path) and are updated along with the live ones
                const bool act = pos > k;
                const float lm = xdiv(act ? a[k] : piv, make_divisor(piv));
                if (act) sA[pos * LD + k] = lm;
                constexpr int q1 = (k + 1) / 4;  // the quad of column k+1 first: the next search only needs that column
                // two columns per instruction (FMUL2 / FFMA2, common.cuh).  Columns <= k of this row are dead data by now
                // (the multiplier has been taken from a[k]), so the pairs that straddle k are updated whole
                {
                    const float4 u = ur4[q1];
                    if (4 * q1 + 1 > k) msub_pair<FUSED>(a[4 * q1], a[4 * q1 + 1], lm, u.x, u.y);
                    msub_pair<FUSED>(a[4 * q1 + 2], a[4 * q1 + 3], lm, u.z, u.w);
                }
                cand = search(std::integral_constant<int, k + 1>{});
#pragma unroll
                for (int q = q1 + 1; q < N / 4; q++) {
                    const float4 u = ur4[q];
                    msub_pair<FUSED>(a[4 * q], a[4 * q + 1], lm, u.x, u.y);
                    msub_pair<FUSED>(a[4 * q + 2], a[4 * q + 3], lm, u.z, u.w);
                }
            }
        });
        fence_async_smem();
        gsync();  // every multiplier is in the slab, P is complete
        if (live) {
            if (r == 0) {
                bulk_s2g(P + b * (long)N, sP, N * 4);
                bulk_commit();
            }
            // L, row by row: thread r supplies column r (multiplier below the diagonal, 1 on it, 0 above)
            float* lg = L + b * (long)(N * N) + r;
#pragma unroll 8
            for (int i = 0; i < N; i++) __stcs(lg + i * N, r < i ? sA[i * LD + r] : (r == i ? 1.0f : 0.0f));
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// shared body: elimination of one matrix `w` (leading dimension ld) by a group of G threads.
// SYNC() must order all G threads; `red` is scratch for the pivot reduction (G/32 candidates).
// ------------------------------------------------------------------------------------------------
template <int G, typename SyncF>
__device__ __forceinline__ void lu_eliminate(float* w, long ld, int n, int t, int pivoting, int* piv_out,
                                             PivotCand* red, SyncF sync) {
    for (int k = 0; k < n - 1; k++) {
        int p = k;
        if (pivoting) {
            PivotCand c{-2.0f, 0x7fffffff};
            for (int i = k + t; i < n; i += G) c = pivot_better(c, PivotCand{pivot_key(w[i * ld + k], i == k), i});
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                PivotCand o{__shfl_xor_sync(0xffffffffu, c.v, off), __shfl_xor_sync(0xffffffffu, c.i, off)};
                c = pivot_better(c, o);
            }
            if constexpr (G > 32) {
                if ((t & 31) == 0) red[t >> 5] = c;
                sync();
                c = red[0];
#pragma unroll
                for (int q = 1; q < G / 32; q++) c = pivot_better(c, red[q]);
                sync();  // red is reused next step
            }
            p = c.i;
        }
        if (t == 0) piv_out[k] = p;
        if (p != k)
            for (int j = k + t; j < n; j += G) {
                float tmp = w[p * ld + j];
                w[p * ld + j] = w[k * ld + j];
                w[k * ld + j] = tmp;
            }
        sync();
        const float diag = w[k * ld + k];
        for (int i = k + 1 + t; i < n; i += G) w[i * ld + k] = xdiv(w[i * ld + k], diag);
        sync();
        {  // trailing update: a warp-wide strip of a row per inner loop (no index division)
            constexpr int GY = G / 32;
            const int tx = t & 31, ty = t >> 5;
            const float* urow = w + k * ld;
            for (int i = k + 1 + ty; i < n; i += GY) {
                float* wrow = w + i * ld;
                const float l = wrow[k];
                for (int j = k + 1 + tx; j < n; j += 32) wrow[j] = xmsub(wrow[j], l, urow[j]);
            }
        }
        sync();
    }
    if (t == 0) piv_out[n - 1] = n - 1;
}

template <int G>
__global__ void __launch_bounds__(256)
lu_group_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ U, int* __restrict__ P, int n,
                long batch, int pivoting) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int kGroups = 256 / G;
    __shared__ PivotCand red_all[8];
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ld = n | 1;
    float* w = smem_f + (size_t)g * (size_t)n * ld;
    PivotCand* red = red_all + g * (G / 32);
    const long nn = (long)n * n;
    auto sync = [g]() { group_sync<G>(g); };
    for (long b = (long)blockIdx.x * kGroups + g; b < batch; b += (long)gridDim.x * kGroups) {
        const float* a = A + b * nn;
        for (int e = t; e < n * n; e += G) {
            int i = e / n, j = e - i * n;
            w[i * ld + j] = a[e];
        }
        sync();
        lu_eliminate<G>(w, ld, n, t, pivoting, P + b * n, red, sync);
        float* l = L + b * nn;
        float* u = U + b * nn;
        for (int e = t; e < n * n; e += G) {
            int i = e / n, j = e - i * n;
            float v = w[i * ld + j];
            l[e] = (j < i) ? v : (j == i ? 1.0f : 0.0f);
            u[e] = (j >= i) ? v : 0.0f;
        }
        sync();
    }
}

__global__ void __launch_bounds__(1024)
lu_global_kernel(const float* __restrict__ A, float* __restrict__ L, float* __restrict__ U, int* __restrict__ P, int n,
                 long batch, int pivoting) {
    __shared__ PivotCand red[32];
    const int t = threadIdx.x;
    const long nn = (long)n * n;
    auto sync = []() { __syncthreads(); };
    for (long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float* a = A + b * nn;
        float* w = U + b * nn;  // eliminate in place in the U buffer, split into L and U afterwards
        float* l = L + b * nn;
        for (long e = t; e < nn; e += 1024) w[e] = a[e];
        __syncthreads();
        lu_eliminate<1024>(w, n, n, t, pivoting, P + b * n, red, sync);
        for (long e = t; e < nn; e += 1024) {
            int i = (int)(e / n), j = (int)(e - (long)i * n);
            float v = w[e];
            l[e] = (j < i) ? v : (j == i ? 1.0f : 0.0f);
            if (j < i) w[e] = 0.0f;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <int N, bool FUSED>
static cudaError_t launch_lu_tpm(const float* A, float* L, float* U, int* P, long batch, int pivoting, cudaStream_t st) {
    constexpr int WARPS = 8;
    constexpr int kSlot = N * N + 4;
    const size_t smem = (size_t)WARPS * 2 * 32 * kSlot * 4 + WARPS * sizeof(uint64_t);
    auto kern = lu_tpm_kernel<N, WARPS, FUSED>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + 31) / 32;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, L, U, P, batch, pivoting);
    count_launch();
    return cudaGetLastError();
}

template <int N, int R, int WARPS, bool FUSED, bool U_DIRECT = false>
static cudaError_t launch_lu_regs(const float* A, float* L, float* U, int* P, long batch, int pivoting, cudaStream_t st) {
    constexpr int MPW = 32 / R;
    constexpr size_t per_warp = (size_t)MPW * ((U_DIRECT ? 1 : 2) * (N * N + 4) + N) * 4;
    static_assert(WARPS * per_warp + WARPS * 8 <= 227 * 1024, "slabs do not fit");
    const size_t smem = WARPS * per_warp + WARPS * sizeof(uint64_t);
    // (one bulk copy per matrix: measured 7-13 % faster here than one part per lane, the opposite of the
    // HBM-bound Cholesky kernels, which prefer many small copies in flight)
    auto kern = pivoting ? lu_regs_kernel<N, R, WARPS, true, U_DIRECT, FUSED> : lu_regs_kernel<N, R, WARPS, false, U_DIRECT, FUSED>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + MPW - 1) / MPW;
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count());
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, L, U, P, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, bool FUSED>
static cudaError_t launch_lu_lane(const float* A, float* L, float* U, int* P, long batch, int pivoting, cudaStream_t st) {
    constexpr int WPM = N / 32;
    constexpr size_t per_mat = (size_t)(N * (N + 1) + 2 * WPM * N + N + 4 * WPM) * 4;
    constexpr int fit = (int)((227 * 1024 - 1024) / (per_mat + 8));
    // registers bound the residency before shared memory does: a row of N floats per lane plus ~38 more
    // (N = 64: 10 matrices per CTA; measured 9 / 10 / 12 matrices: 16.6 / 17.7 / 17.0 M fact/s)
    constexpr int by_regs = 65536 / ((N + 38) * N);
    constexpr int MATS = fit * N > 1024 ? 1024 / N : (fit < by_regs ? fit : by_regs);
    static_assert(MATS >= 1 && (WPM == 1 || MATS <= 15), "named barrier ids 1..15");
    const size_t smem = MATS * per_mat + MATS * sizeof(uint64_t);
    auto kern = pivoting ? lu_lane_kernel<N, MATS, true, FUSED> : lu_lane_kernel<N, MATS, false, FUSED>;
    // (LOCKSTEP measured -18 % at N = 64 and -1 % at N = 128: the per-matrix barrier stays)
    if (N >= 64 && pivoting && option(HLSD_OPT_LOCKSTEP) == 1) kern = lu_lane_kernel<N, MATS, true, FUSED, (N >= 64)>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    long grid = std::min<long>((batch + MATS - 1) / MATS, (long)sm_count());
    kern<<<(unsigned)grid, MATS * N, smem, st>>>(A, L, U, P, batch);
    count_launch();
    return cudaGetLastError();
}

template <int G>
static cudaError_t launch_lu_group(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting,
                                   cudaStream_t st) {
    constexpr int kGroups = 256 / G;
    const size_t smem = (size_t)n * (n | 1) * 4 * kGroups;
    auto kern = lu_group_kernel<G>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((batch + kGroups - 1) / kGroups, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, 256, smem, st>>>(A, L, U, P, n, batch, pivoting);
    count_launch();
    return cudaGetLastError();
}

template <bool FUSED>
static cudaError_t lu_dispatch(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, int path,
                               cudaStream_t st) {
    const bool aligned = (((uintptr_t)A | (uintptr_t)L | (uintptr_t)U | (uintptr_t)P) & 15) == 0;
    if (path == HLSD_PATH_AUTO || path == HLSD_PATH_TPM) {
        if (aligned && n == 4) return launch_lu_tpm<4, FUSED>(A, L, U, P, batch, pivoting, st);
        // (thread-per-matrix at N = 8 - 148 registers, pivot exchanges as predicated swaps - measured 4.5 G against 5.2 G fact/s)
        if (aligned && n == 8) return launch_lu_regs<8, 2, 20, FUSED>(A, L, U, P, batch, pivoting, st);
        // (eight lanes per matrix at N = 16: 0.88 G fact/s with 24 warps, 0.81 G with 16, against 1.03 G for four lanes)
        if (aligned && n == 16) return launch_lu_regs<16, 4, 12, FUSED, true>(A, L, U, P, batch, pivoting, st);  // U rows straight out: +10 %
        if (aligned && n == 32) {
            // measured (round 2, batch 259441): lane-per-row kernel 227 M fact/s; the rows-in-registers kernel
            // (lu_regs_kernel<32, 16>) managed 116 M (76 M after the lane kernel's last changes) and is no longer built
            return launch_lu_lane<32, FUSED>(A, L, U, P, batch, pivoting, st);
        }
        if (aligned && n == 64) return launch_lu_lane<64, FUSED>(A, L, U, P, batch, pivoting, st);
        if (aligned && n == 128) return launch_lu_lane<128, FUSED>(A, L, U, P, batch, pivoting, st);
        if (path == HLSD_PATH_TPM) return cudaErrorInvalidValue;
    }
    const size_t per = (size_t)n * (n | 1) * 4;
    const size_t cap = 227 * 1024 - 1024;
    if ((path == HLSD_PATH_AUTO && per <= cap) || path == HLSD_PATH_GROUP) {
        if (per > cap) return cudaErrorInvalidValue;
        if (n <= 48 && per * 8 <= cap) return launch_lu_group<32>(A, L, U, P, n, batch, pivoting, st);
        if (n <= 96 && per * 2 <= cap) return launch_lu_group<128>(A, L, U, P, n, batch, pivoting, st);
        return launch_lu_group<256>(A, L, U, P, n, batch, pivoting, st);
    }
    long grid = std::min<long>(batch, (long)sm_count() * 2);
    lu_global_kernel<<<(unsigned)grid, 1024, 0, st>>>(A, L, U, P, n, batch, pivoting);
    count_launch();
    return cudaGetLastError();
}

cudaError_t lu_exact(const float* A, float* L, float* U, int* P, int n, long batch, int pivoting, int path, int fused,
                     cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    ProfScope ps(HLSD_PROF_EXACT, st);
    return fused ? lu_dispatch<true>(A, L, U, P, n, batch, pivoting, path, st)
                 : lu_dispatch<false>(A, L, U, P, n, batch, pivoting, path, st);
}

}  // namespace hlsd
