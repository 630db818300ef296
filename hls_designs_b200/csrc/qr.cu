// Batched QR by Givens rotations, the reference's `int top(A, Q, R)` (QR/qr_v*/model4x4/qr.h:22-24).
//
//   variant 0  qr_v2.1: per column c, rows i = ROWS-1 .. c+1, rotate adjacent rows (i-1, i); a
//              rotation is skipped (c = 1, s = 0, still applied) when A[i] == 0
//              (qr_v2.1/template/T_qr.cpp:63-148 angle generation, :150-223 row rotation,
//              :225-300 Q accumulation).  A true QR: A = QR.
//   variant 1  qr_v1.1 / qr_v1.2: rows paired (i-2, i) (column 0 finishes with (0, 1)), skip test
//              |A[i]| < 1e-6 (qr_v1.1/model4x4/qr.cpp:45-127, 131-226, 310-404).  Reproduced for
//              element-wise parity only: those designs do not satisfy A = QR.
//
// The systolic arrays run the rotations of different columns concurrently; so does variant 0 here:
// rotation (i, c) only conflicts with rotations that share one of its rows, and the schedule
// time(i, c) = (ROWS-1-i) + 2c orders every conflicting pair as the sequential loop does, so each
// element sees the same operations in the same order and the result is bit-identical.
//
// Kernels:
//   qr_tpm_kernel<N>     thread-per-matrix in registers, square N in {4, 8}.
//   qr_wcols_kernel<N,K> wavefront schedule in registers, K columns of R and K rows of Q per lane, square
//                        N in {16, 32}, qr_v2.1 arithmetic.
//   qr_cols_kernel<N>    sequential schedule, lane per column (qr_v1.x arithmetic at N = 16).
//   qr_qreg_kernel<N>    wavefront with R in shared memory and the rows of Q in registers, square N in {64, 128},
//                        qr_v2.1 arithmetic (QR 64: 4.2 M -> 5.2 M fact/s, QR 128: 402 k -> 544 k against qr_wave_kernel).
//   qr_wave_kernel<G>    G threads per matrix, R and Q resident in shared memory, wavefront schedule
//                        (variant 0) or sequential schedule (variant 1).
//   qr_global_kernel     one CTA per matrix, in place in global memory (any size; exact; slow).
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

// angle generation (qrf_mag + the two divisions), QR/qr_v2.1/model4x4/qr.cpp:36-43 and :96-100
__device__ __forceinline__ void givens_make(float& lo, float& hi, bool skip, float& c, float& s) {
    if (skip) {
        c = 1.0f;
        s = 0.0f;
    } else {
        float mag = xsqrt(xadd(xmul(hi, hi), xmul(lo, lo)));
        c = xdiv(lo, mag);
        s = xdiv(hi, mag);
        lo = mag;
        hi = 0.0f;
    }
}
template <int VARIANT>
__device__ __forceinline__ bool givens_skip(float hi) {
    if constexpr (VARIANT == 0) return hi == 0.0f;
    else return fabsf(hi) < 1e-6;  // double comparison, as in the reference (`1e-6` is a double literal)
}
// qr_v2.1 angle generation with one reciprocal for the two divisions (make_divisor / fast_quotient, common.cuh);
// a skipped rotation (hi == 0) computes on (1, 1) so that it stays on the fast paths, and is then discarded
template <bool FUSED>
__device__ __forceinline__ float givens_mag(float x, float y) {
    if constexpr (FUSED) return xsqrt(__fmaf_rn(x, x, xmul(y, y)));
    else return xsqrt(xadd(xmul(x, x), xmul(y, y)));
}
template <bool FUSED = false>
__device__ __forceinline__ void givens_make_fast(float& lo, float& hi, float& c, float& s) {
    const bool skip = givens_skip<0>(hi);
    const float x = skip ? 1.0f : hi, y = skip ? 1.0f : lo;
    const float mag = givens_mag<FUSED>(x, y);
    const SharedDivisor dv = make_divisor(mag);
    float cq, sq;
    if (dv.ok && div_in_range(x) && div_in_range(y)) {
        cq = fast_quotient(y, dv);
        sq = fast_quotient(x, dv);
    } else {
        cq = xdiv(y, mag);
        sq = xdiv(x, mag);
    }
    c = skip ? 1.0f : cq;
    s = skip ? 0.0f : sq;
    lo = skip ? lo : mag;
    hi = skip ? hi : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// thread-per-matrix (registers): square N x N
// ------------------------------------------------------------------------------------------------
template <int N, int VARIANT, int WARPS, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, 1)
qr_tpm_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, long batch) {
    constexpr int kSlot = N * N + 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * 2 * 32 * kSlot * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* slotR = slabs + ((size_t)warp * 2 * 32 + lane) * kSlot;
    float* slotQ = slotR + 32 * kSlot;
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long num_slabs = (batch + 31) / 32;
    uint32_t parity = 0;
    for (long sl = (long)blockIdx.x * WARPS + warp; sl < num_slabs; sl += (long)gridDim.x * WARPS) {
        const long m0 = sl * 32;
        const int valid = (int)min(32L, batch - m0);
        constexpr bool kCoalesced = N <= 4;  // 64-byte matrices: cooperative coalesced copies beat one bulk copy per matrix
        constexpr uint32_t kMatBytes = N * N * 4;
        float4* slabR = reinterpret_cast<float4*>(slabs + (size_t)warp * 2 * 32 * kSlot);
        float4* slabQ = slabR + 32 * (kSlot / 4);
        if constexpr (kCoalesced) {
            __syncwarp();
            slab_g2s<N * N / 4>(slabR, reinterpret_cast<const float4*>(A + m0 * (long)(N * N)), valid, lane);
            __syncwarp();
        } else {
            bulk_wait_read<0>();
            if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * kMatBytes);
            __syncwarp();
            if (lane < valid) bulk_g2s(slotR, A + (m0 + lane) * (long)(N * N), kMatBytes, bar);
            mbar_wait(bar, parity);
            parity ^= 1;
        }
        if (lane < valid) {
            float r[N][N], q[N][N];
            const float4* s4 = reinterpret_cast<const float4*>(slotR);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int v = 0; v < N / 4; v++) {
                    float4 x = s4[i * (N / 4) + v];
                    r[i][4 * v] = x.x, r[i][4 * v + 1] = x.y, r[i][4 * v + 2] = x.z, r[i][4 * v + 3] = x.w;
                }
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = 0; j < N; j++) q[i][j] = (i == j) ? 1.0f : 0.0f;
#pragma unroll
            for (int col = 0; col < N; col++) {
#pragma unroll
                for (int i = N - 1; i > col; i--) {
                    const int lo = (VARIANT == 0) ? i - 1 : ((col == 0 && i < 2) ? i - 1 : i - 2);
                    float c, s;
                    givens_make(r[lo][col], r[i][col], givens_skip<VARIANT>(r[i][col]), c, s);
#pragma unroll
                    for (int j = col + 1; j < N; j++) givens_apply_t<FUSED>(c, s, r[lo][j], r[i][j]);
#pragma unroll
                    for (int k = 0; k < N; k++) givens_apply_t<FUSED>(c, s, q[k][lo], q[k][i]);
                }
            }
            float4* r4 = reinterpret_cast<float4*>(slotR);
            float4* q4 = reinterpret_cast<float4*>(slotQ);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int v = 0; v < N / 4; v++) {
                    r4[i * (N / 4) + v] = make_float4(r[i][4 * v], r[i][4 * v + 1], r[i][4 * v + 2], r[i][4 * v + 3]);
                    q4[i * (N / 4) + v] = make_float4(q[i][4 * v], q[i][4 * v + 1], q[i][4 * v + 2], q[i][4 * v + 3]);
                }
            if constexpr (!kCoalesced) {
                fence_async_smem();
                bulk_s2g(R + (m0 + lane) * (long)(N * N), slotR, kMatBytes);
                bulk_s2g(Q + (m0 + lane) * (long)(N * N), slotQ, kMatBytes);
                bulk_commit();
            }
        }
        if constexpr (kCoalesced) {
            __syncwarp();
            slab_s2g<N * N / 4>(reinterpret_cast<float4*>(R + m0 * (long)(N * N)), slabR, valid, lane);
            slab_s2g<N * N / 4>(reinterpret_cast<float4*>(Q + m0 * (long)(N * N)), slabQ, valid, lane);
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// N lanes per matrix (square N = 16, 32): lane j keeps COLUMN j of R and ROW j of Q in registers.
// A rotation of rows (lo, i) then touches two registers of every lane's R column and two of its Q
// row - no data moves between lanes except the generating pair: the lane that owns column c
// broadcasts it by shuffle (the systolic arrays stream (c, s) from the diagonal PE to the PEs on
// its right, qr_v2.1/template/T_qr.cpp PE1 -> PE2; here every lane derives the same (c, s)).
// Rotations run in the reference's sequential order (column by column, bottom up), so every element
// sees the same operations in the same order: bit-identical.  Row indices are compile-time (the
// row loop is unrolled and predicated on i > c), the column loop is a run-time loop.
// ------------------------------------------------------------------------------------------------
template <int N, int VARIANT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
qr_cols_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, long batch) {
    constexpr int MPW = 32 / N;  // matrices per warp
    constexpr int kSlot = N * N + 4;
    constexpr uint32_t kRowBytes = N * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * MPW * 2 * kSlot * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = lane % N, m = lane / N, gbase = lane - j;
    float* sR = slabs + ((size_t)warp * MPW + m) * 2 * kSlot;
    float* sQ = sR + kSlot;
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long num_slabs = (batch + MPW - 1) / MPW;
    uint32_t parity = 0;
    for (long sl = (long)blockIdx.x * WARPS + warp; sl < num_slabs; sl += (long)gridDim.x * WARPS) {
        const long m0 = sl * MPW;
        const int valid = (int)min((long)MPW, batch - m0);
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * N * kRowBytes);
        __syncwarp();
        if (m < valid) bulk_g2s(sR + j * N, A + (m0 + m) * (long)(N * N) + j * N, kRowBytes, bar);
        mbar_wait(bar, parity);
        parity ^= 1;
        float r[N], q[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            r[i] = sR[i * N + j];  // column j of A
            q[i] = (i == j) ? 1.0f : 0.0f;  // row j of Q = e_j
        }
        for (int c = 0; c < N - 1; c++) {
            static_for<1, N>([&](auto ic) {
                constexpr int i = N - decltype(ic)::value;  // N-1 down to 1
                constexpr int lo = (VARIANT == 0) ? i - 1 : (i == 1 ? 0 : i - 2);
                if (i > c) {  // warp-uniform
                    // the generating pair lives in the lane that owns column c: broadcast it and let every lane
                    // derive the same angle (all lanes then run the fast paths of sqrt / division together,
                    // instead of one lane computing on real data and the others on unrelated values)
                    float hi_v = __shfl_sync(0xffffffffu, r[i], gbase + c);
                    float lo_v = __shfl_sync(0xffffffffu, r[lo], gbase + c);
                    const bool skip = givens_skip<VARIANT>(hi_v);
                    float cs, sn;
                    givens_make(lo_v, hi_v, skip, cs, sn);
                    if (j == c) {
                        r[lo] = lo_v;  // mag (or unchanged when skipped)
                        r[i] = hi_v;   // 0   (or unchanged when skipped)
                    } else if (j > c) {
                        givens_apply(cs, sn, r[lo], r[i]);
                    }
                    givens_apply(cs, sn, q[lo], q[i]);
                }
            });
        }
        __syncwarp();  // every lane has read its column from the slab
        if (m < valid) {
#pragma unroll
            for (int i = 0; i < N; i++) sR[i * N + j] = r[i];
            float4* q4 = reinterpret_cast<float4*>(sQ + j * N);
#pragma unroll
            for (int v = 0; v < N / 4; v++) q4[v] = make_float4(q[4 * v], q[4 * v + 1], q[4 * v + 2], q[4 * v + 3]);
            fence_async_smem();
        }
        __syncwarp();
        if (m < valid) {
            bulk_s2g(R + (m0 + m) * (long)(N * N) + j * N, sR + j * N, kRowBytes);
            bulk_s2g(Q + (m0 + m) * (long)(N * N) + j * N, sQ + j * N, kRowBytes);
            bulk_commit();
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// Wavefront in registers (qr_v2.1 arithmetic, square N = 16, 32): C = N / K lanes per matrix, lane l keeps
// columns l, l + C, ... of R and rows l, l + C, ... of Q (K of each).  The schedule is the wavefront
// time(i, c) = (N-1-i) + 2c of qr_factor below, fully unrolled, so every row index is a compile-time
// constant.  The rotations of one step have consecutive columns c, i.e. distinct owner lanes c % C: each
// owner generates its own angle from its own column, all of them in the same instructions (the systolic
// arrays' diagonal PEs, qr_v2.1/template/T_qr.cpp:63-148, work side by side in the same way), then every
// (c, s) is broadcast by shuffle and applied by all lanes of the group to their columns of R to the right
// of c and to their rows of Q.  Per-element operation order is the sequential loop's: bit-identical.
// ------------------------------------------------------------------------------------------------
template <int N, int K, int WARPS, bool FUSED, bool LOCKSTEP = false>  // LOCKSTEP: see chol_rows_kernel (chol.cu)
__global__ void __launch_bounds__(WARPS * 32, 1)
qr_wcols_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, long batch) {
    constexpr int C = N / K;     // lanes per matrix
    constexpr int MPW = 32 / C;  // matrices per warp
    constexpr int kSlot = N * N + 4;
    constexpr uint32_t kMatBytes = N * N * 4, kPart = kMatBytes / C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slabs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * MPW * 2 * kSlot * 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = lane % C, m = lane / C, gbase = lane - l;
    float* sR = slabs + ((size_t)warp * MPW + m) * 2 * kSlot;
    float* sQ = sR + kSlot;
    uint64_t* bar = bars + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long num_slabs = (batch + MPW - 1) / MPW;
    uint32_t parity = 0;
    for (long sl0 = (long)blockIdx.x * WARPS; sl0 < num_slabs; sl0 += (long)gridDim.x * WARPS) {
        const long sl = sl0 + warp;
        if (!LOCKSTEP && sl >= num_slabs) break;
        const long m0 = sl * MPW;
        const int valid = sl < num_slabs ? (int)min((long)MPW, batch - m0) : 0;
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * kMatBytes);
        __syncwarp();
        if (m < valid) bulk_g2s(sR + l * (kPart / 4), A + (m0 + m) * (long)(N * N) + l * (kPart / 4), kPart, bar);
        mbar_wait(bar, parity);
        parity ^= 1;
        float rr[K][N], qq[K][N];
#pragma unroll
        for (int kk = 0; kk < K; kk++)
#pragma unroll
            for (int i = 0; i < N; i++) {
                rr[kk][i] = sR[i * N + kk * C + l];                 // column kk*C + l of A
                qq[kk][i] = (i == kk * C + l) ? 1.0f : 0.0f;        // row kk*C + l of the identity
            }
        constexpr int kLastStep = (N - 1) + 2 * (N - 2);
        static_for<0, kLastStep + 1>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            constexpr int c_lo = (s - N + 2) > 0 ? (s - N + 2) : 0;
            constexpr int c_hi = (s / 2) < (N - 2) ? (s / 2) : (N - 2);
            constexpr int nrot = c_hi - c_lo + 1;
            if constexpr (LOCKSTEP && nrot > 0) __syncthreads();
            if constexpr (nrot > 0) {
                static_for<0, (nrot + C - 1) / C>([&](auto chc) {
                    constexpr int cb = c_lo + decltype(chc)::value * C;
                    constexpr int cnt = (c_hi - cb + 1) < C ? (c_hi - cb + 1) : C;
                    // every owner lane picks the generating pair of its own rotation (others: a harmless (1, 1))
                    float lo = 1.0f, hi = 1.0f;
                    static_for<0, cnt>([&](auto ec) {
                        constexpr int c = cb + decltype(ec)::value, i = N - 1 - s + 2 * c;
                        if (l == c % C) {
                            lo = rr[c / C][i - 1];
                            hi = rr[c / C][i];
                        }
                    });
                    // angle generation, all owners at once; a skipped rotation (hi == 0) computes on (1, 1)
                    const bool skip = givens_skip<0>(hi);
                    const float x = skip ? 1.0f : hi, y = skip ? 1.0f : lo;
                    const float mag = givens_mag<FUSED>(x, y);
                    const SharedDivisor dv = make_divisor(mag);
                    float cs, sn;
                    if (dv.ok && div_in_range(x) && div_in_range(y)) {
                        cs = fast_quotient(y, dv);
                        sn = fast_quotient(x, dv);
                    } else {
                        cs = xdiv(y, mag);
                        sn = xdiv(x, mag);
                    }
                    cs = skip ? 1.0f : cs;
                    sn = skip ? 0.0f : sn;
                    const float new_lo = skip ? lo : mag, new_hi = skip ? hi : 0.0f;
                    static_for<0, cnt>([&](auto ec) {
                        constexpr int c = cb + decltype(ec)::value, i = N - 1 - s + 2 * c;
                        constexpr int kc = c / C, lc = c % C;
                        const float cc = __shfl_sync(0xffffffffu, cs, gbase + lc);
                        const float ss = __shfl_sync(0xffffffffu, sn, gbase + lc);
                        // R: column kc*C + l takes the rotation when l > lc, is the generating column when l == lc;
                        // columns of the later groups always take it, earlier ones never
                        {
                            const float a0 = rr[kc][i - 1], b0 = rr[kc][i];
                            float a1 = a0, b1 = b0;
                            givens_apply_t<FUSED>(cc, ss, a1, b1);
                            rr[kc][i - 1] = (l == lc) ? new_lo : (l > lc ? a1 : a0);
                            rr[kc][i] = (l == lc) ? new_hi : (l > lc ? b1 : b0);
                        }
                        static_for<kc + 1, K>([&](auto kkc) {
                            constexpr int kk = decltype(kkc)::value;
                            givens_apply_t<FUSED>(cc, ss, rr[kk][i - 1], rr[kk][i]);
                        });
                        static_for<0, K>([&](auto kkc) {
                            constexpr int kk = decltype(kkc)::value;
                            givens_apply_t<FUSED>(cc, ss, qq[kk][i - 1], qq[kk][i]);
                        });
                    });
                });
            }
        });
        __syncwarp();  // every lane has read its columns from the slab
        if (m < valid) {
#pragma unroll
            for (int kk = 0; kk < K; kk++) {
#pragma unroll
                for (int i = 0; i < N; i++) sR[i * N + kk * C + l] = rr[kk][i];
                float4* q4 = reinterpret_cast<float4*>(sQ + (kk * C + l) * N);
#pragma unroll
                for (int v = 0; v < N / 4; v++)
                    q4[v] = make_float4(qq[kk][4 * v], qq[kk][4 * v + 1], qq[kk][4 * v + 2], qq[kk][4 * v + 3]);
            }
            fence_async_smem();
        }
        __syncwarp();
        if (m < valid) {
            bulk_s2g(R + (m0 + m) * (long)(N * N) + l * (kPart / 4), sR + l * (kPart / 4), kPart);
            bulk_s2g(Q + (m0 + m) * (long)(N * N) + l * (kPart / 4), sQ + l * (kPart / 4), kPart);
            bulk_commit();
        }
    }
    bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// shared body: factor one matrix held in r (rows x ldr) / q (rows x ldq) with G threads.
// cs: scratch for 2 * (rows/2 + 1) floats (the (c, s) pairs of one wavefront).
// ------------------------------------------------------------------------------------------------
template <int G, int VARIANT, typename SyncF>
__device__ __forceinline__ void qr_factor(float* r, long ldr, float* q, long ldq, int rows, int cols, int t, float* cs,
                                          SyncF sync) {
    if constexpr (VARIANT == 0) {
        const int last_t = (rows - 1) + 2 * (cols - 1);  // generous bound; empty steps are skipped
        for (int step = 0; step <= last_t; step++) {
            // rotations of this step: columns c with i = rows-1-step+2c, c < i <= rows-1, c < cols
            int c_lo = max(0, step - rows + 2);
            int c_hi = min(cols - 1, step / 2);
            const int nrot = c_hi - c_lo + 1;
            if (nrot <= 0) continue;
            for (int k = t; k < nrot; k += G) {
                const int c = c_lo + k, i = rows - 1 - step + 2 * c;
                float lo = r[(long)(i - 1) * ldr + c], hi = r[(long)i * ldr + c];
                float cc, ss;
                const bool skip = givens_skip<0>(hi);
                givens_make(lo, hi, skip, cc, ss);
                r[(long)(i - 1) * ldr + c] = lo;
                r[(long)i * ldr + c] = hi;
                cs[2 * k] = cc;
                cs[2 * k + 1] = ss;
            }
            sync();
            // apply: thread (ty, tx) takes rotations k = ty, ty + GY, ... and elements x = tx, tx + GX, ...
            // (columns of R to the right of c, then rows of Q); rotations of one step touch disjoint rows.
            // Everything that does not depend on the rotation is hoisted: an element x is either column x of
            // R (stride ldr between rows) or row x - cols of Q (stride 1 between columns).
            const int width = cols + rows;
            constexpr int GX = G > 256 ? 256 : G, GY = G / GX;
            const int tx = t % GX, ty = t / GX;
            const int i0 = rows - 1 - step + 2 * c_lo;  // row index of rotation k: i0 + 2 k
            for (int x = tx; x < width; x += GX) {
                const bool is_r = x < cols;
                float* base = is_r ? r + x : q + (long)(x - cols) * ldq;
                const int stride = is_r ? (int)ldr : 1;
                const int kmax = is_r ? min(nrot, x - c_lo) : nrot;  // column x of R only takes rotations with c < x
#pragma unroll 4
                for (int k = ty; k < kmax; k += GY) {
                    const float cc = cs[2 * k], ss = cs[2 * k + 1];
                    float* p = base + (i0 + 2 * k - 1) * stride;
                    float lo = p[0], hi = p[stride];
                    givens_apply(cc, ss, lo, hi);
                    p[0] = lo;
                    p[stride] = hi;
                }
            }
            sync();
        }
    } else {
        for (int col = 0; col < cols; col++)
            for (int i = rows - 1; i > col; i--) {
                const int lo_row = (col == 0 && i < 2) ? i - 1 : i - 2;
                float lo = r[(long)lo_row * ldr + col], hi = r[(long)i * ldr + col];
                float cc, ss;
                givens_make(lo, hi, givens_skip<1>(hi), cc, ss);
                sync();  // every thread has read the generating pair
                if (t == 0) {
                    r[(long)lo_row * ldr + col] = lo;
                    r[(long)i * ldr + col] = hi;
                }
                const int width = cols + rows;
                for (int x = t; x < width; x += G) {
                    if (x < cols) {
                        if (x > col) givens_apply(cc, ss, r[(long)lo_row * ldr + x], r[(long)i * ldr + x]);
                    } else {
                        const int row = x - cols;
                        givens_apply(cc, ss, q[(long)row * ldq + lo_row], q[(long)row * ldq + i]);
                    }
                }
                sync();
            }
    }
}

template <int G, int VARIANT>
__global__ void __launch_bounds__(G > 256 ? G : 256)
qr_wave_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, int rows, int cols,
               long batch) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int kBlock = G > 256 ? G : 256;
    constexpr int kGroups = kBlock / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ldr = cols | 1, ldq = rows | 1;
    const size_t per = (size_t)rows * ldr + (size_t)rows * ldq + 2 * (rows / 2 + 2);
    float* r = smem_f + g * per;
    float* q = r + (size_t)rows * ldr;
    float* cs = q + (size_t)rows * ldq;
    auto sync = [g]() {
        if constexpr (G > 256) __syncthreads();
        else group_sync<G>(g);
    };
    const long na = (long)rows * cols, nq = (long)rows * rows;
    for (long b = (long)blockIdx.x * kGroups + g; b < batch; b += (long)gridDim.x * kGroups) {
        const float* a = A + b * na;
        for (int e = t; e < rows * cols; e += G) {
            int i = e / cols, j = e - i * cols;
            r[i * ldr + j] = a[e];
        }
        for (int e = t; e < rows * rows; e += G) {
            int i = e / rows, j = e - i * rows;
            q[i * ldq + j] = (i == j) ? 1.0f : 0.0f;
        }
        sync();
        qr_factor<G, VARIANT>(r, ldr, q, ldq, rows, cols, t, cs, sync);
        float* ro = R + b * na;
        float* qo = Q + b * nq;
        for (int e = t; e < rows * cols; e += G) {
            int i = e / cols, j = e - i * cols;
            ro[e] = r[i * ldr + j];
        }
        for (int e = t; e < rows * rows; e += G) {
            int i = e / rows, j = e - i * rows;
            qo[e] = q[i * ldq + j];
        }
        sync();
    }
}

// ------------------------------------------------------------------------------------------------
// Wavefront with Q in registers (qr_v2.1 arithmetic, square N = 64, 128): one CTA of RT + N threads per matrix.
// Threads 0..RT-1 work on R in shared memory exactly like qr_factor (angle generation, columns right of c);
// thread RT + j keeps ROW j of Q - two thirds of all rotation work - in N registers.  A rotation of a step
// touches rows (hi-1, hi) with hi = N-1-step + 2c, so within one step all pairs have the same alignment:
// pair slot p = rows (2p + par - 1, 2p + par), par = parity of N-1-step, and the active slots are the
// contiguous range p0 + c_lo .. p0 + c_hi.  The slot loop is unrolled (register indices are constants),
// guarded per group of 8 and per slot by warp-uniform range tests; slot p finds its (c, s) at a constant
// offset from a per-step base pointer.  Per rotation and Q row: one 8-byte shared-memory load and the six
// floating-point operations, instead of two loads, two stores and address arithmetic on top.
// ------------------------------------------------------------------------------------------------
// UNI: ONE set of N threads per matrix does both jobs (thread t: the R work of R-thread t, then row t of Q), which
// halves the threads and registers a matrix occupies: two (128 x 128) or more CTAs share an SM, and one matrix's
// angle generation and barrier waits are filled by another matrix's rotations.  (ncu on the two-role kernel at
// 128 x 128, one CTA per SM: issue slots 45 % busy, `barrier` + `wait` the top stalls.)
// (three CTAs per SM at N = 128 - a 168-register cap - spill the Q row in the slot loop: 683 k -> 299 k fact/s)
template <int N, int RT, bool FUSED, bool UNI = false>
__global__ void __launch_bounds__(UNI ? N : RT + N, UNI ? (N == 64 ? 8 : 2) : (N == 64 ? 4 : 1))
qr_qreg_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, long batch) {
    static_assert(!UNI || RT == N, "unified threads: one thread per R column group and Q row");
    constexpr int NT = UNI ? N : RT + N;  // threads per CTA
    extern __shared__ __align__(16) float smem_f[];
    constexpr int ldr = N | 1;
    float* r = smem_f;                                   // N x ldr
    // (c, s) of the step's rotations, with 8 entries of slack in front and 16 behind (partially active slot groups load
    // unconditionally and discard); N * ldr is even, so the buffer is 8-byte aligned
    float2* cs2 = reinterpret_cast<float2*>(r + N * ldr) + 8;
    const int t = threadIdx.x;
    const bool q_side = UNI || t >= RT;  // threads 0..RT-1 work on R, thread RT + j (UNI: thread j) holds row j of Q
    const bool r_side = UNI || t < RT;
    const int qrow = UNI ? t : t - RT;
    for (long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float* a = A + b * (long)(N * N);
        for (int e = t; e < N * N; e += NT) r[(e / N) * ldr + (e % N)] = a[e];
        float qq[N];
#pragma unroll
        for (int i = 0; i < N; i++) qq[i] = (q_side && i == qrow) ? 1.0f : 0.0f;
        __syncthreads();
        constexpr int last_t = (N - 1) + 2 * (N - 1);
        for (int step = 0; step <= last_t; step++) {
            const int c_lo = max(0, step - N + 2);
            const int c_hi = min(N - 1, step / 2);
            const int nrot = c_hi - c_lo + 1;
            if (nrot <= 0) continue;
            const int i0 = N - 1 - step + 2 * c_lo;  // row index of rotation k: i0 + 2 k
            if (r_side) {
                for (int k = t; k < nrot; k += RT) {
                    const int c = c_lo + k, i = i0 + 2 * k;
                    float lo = r[(i - 1) * ldr + c], hi = r[i * ldr + c];
                    float cc, ss;
                    givens_make_fast<FUSED>(lo, hi, cc, ss);
                    r[(i - 1) * ldr + c] = lo;
                    r[i * ldr + c] = hi;
                    cs2[k] = make_float2(cc, ss);
                }
            }
            __syncthreads();
            if (r_side) {
                // R: columns x = tx, tx + 32, ... take the rotations with c < x; rotations k = ty, ty + GY, ...
                constexpr int GY = RT / 32;
                const int tx = t & 31, ty = t >> 5;
                // sixteen (N = 128) independent rotations per turn and thread: four rotations on each of the thread's
                // columns; all loads, then the arithmetic, then all stores (the rotations of a step touch disjoint
                // rows, which the compiler cannot know)
                constexpr int XC = N / 32;
                float* base = r + tx + (i0 - 1) * ldr;
                for (int k = ty; k < nrot; k += 4 * GY) {
                    float lo[XC][4], hi[XC][4];
                    float2 w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) w[j] = cs2[k + j * GY];  // (slack entries behind the buffer)
#pragma unroll
                    for (int xc = 0; xc < XC; xc++) {
                        const int kmax = min(nrot, tx + 32 * xc - c_lo);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int kk = k + j * GY;
                            if (kk < kmax) {
                                lo[xc][j] = base[2 * kk * ldr + 32 * xc];
                                hi[xc][j] = base[(2 * kk + 1) * ldr + 32 * xc];
                            }
                        }
                    }
#pragma unroll
                    for (int xc = 0; xc < XC; xc++) {
                        const int kmax = min(nrot, tx + 32 * xc - c_lo);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (k + j * GY < kmax) givens_apply_t<FUSED>(w[j].x, w[j].y, lo[xc][j], hi[xc][j]);
                    }
#pragma unroll
                    for (int xc = 0; xc < XC; xc++) {
                        const int kmax = min(nrot, tx + 32 * xc - c_lo);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int kk = k + j * GY;
                            if (kk < kmax) {
                                base[2 * kk * ldr + 32 * xc] = lo[xc][j];
                                base[(2 * kk + 1) * ldr + 32 * xc] = hi[xc][j];
                            }
                        }
                    }
                }
            }
            if (q_side) {
                const int bb = N - 1 - step;  // hi row of rotation c: bb + 2 c
                const int p0 = bb >> 1;       // floor(bb / 2): slot of rotation c is p0 + c
                const int plo = p0 + c_lo, phi = p0 + c_hi;
                const float2* csp = cs2 - plo;  // slot p reads csp[p]
                auto run = [&](auto parc) {
                    constexpr int PAR = decltype(parc)::value;
                    static_for<0, N / 16>([&](auto gc) {
                        constexpr int g = decltype(gc)::value;
                        if (8 * g >= plo && 8 * g + 7 <= phi) {
                            // all eight slots active: straight-line code, the eight loads go out together
                            static_for<0, 8>([&](auto ec) {
                                constexpr int p = 8 * g + decltype(ec)::value;
                                constexpr int lo = 2 * p + PAR - 1, hi = 2 * p + PAR;
                                if constexpr (lo >= 0) {
                                    const float2 w = csp[p];
                                    givens_apply_t<FUSED>(w.x, w.y, qq[lo], qq[hi]);
                                }
                            });
                        } else if (8 * g + 7 >= plo && 8 * g <= phi) {
                            // partially active group: unconditional loads (slack entries around the buffer), results selected
                            static_for<0, 8>([&](auto ec) {
                                constexpr int p = 8 * g + decltype(ec)::value;
                                constexpr int lo = 2 * p + PAR - 1, hi = 2 * p + PAR;
                                if constexpr (lo >= 0) {
                                    const bool on = p >= plo && p <= phi;
                                    const float2 w = csp[p];
                                    float x0 = qq[lo], x1 = qq[hi];
                                    givens_apply_t<FUSED>(w.x, w.y, x0, x1);
                                    qq[lo] = on ? x0 : qq[lo];
                                    qq[hi] = on ? x1 : qq[hi];
                                }
                            });
                        }
                    });
                };
                if (bb & 1) run(std::integral_constant<int, 1>{});
                else run(std::integral_constant<int, 0>{});
            }
            __syncthreads();
        }
        float* ro = R + b * (long)(N * N);
        for (int e = t; e < N * N; e += NT) ro[e] = r[(e / N) * ldr + (e % N)];
        if (q_side) {
            float4* q4 = reinterpret_cast<float4*>(Q + b * (long)(N * N) + (long)qrow * N);
#pragma unroll
            for (int v = 0; v < N / 4; v++) q4[v] = make_float4(qq[4 * v], qq[4 * v + 1], qq[4 * v + 2], qq[4 * v + 3]);
        }
        __syncthreads();
    }
}

// (Round 2 experiment, removed: the same wavefront with the COLUMNS of R in registers as well - thread x generates its own
// rotation from a carried magnitude and a captured neighbour value, one barrier per step, 2 N threads and 254 registers
// per matrix.  Bit-identical, but 344-383 k fact/s at 128 x 128 against 683 k for the kernel above: one matrix per SM
// leaves 8 warps to hide the instruction fetch of the unrolled slot code (`no_instruction` 2.3 warps per issue).)

template <int VARIANT>
__global__ void __launch_bounds__(1024)
qr_global_kernel(const float* __restrict__ A, float* __restrict__ Q, float* __restrict__ R, int rows, int cols,
                 long batch) {
    extern __shared__ __align__(16) float cs[];
    const int t = threadIdx.x;
    auto sync = []() { __syncthreads(); };
    const long na = (long)rows * cols, nq = (long)rows * rows;
    for (long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float* a = A + b * na;
        float* r = R + b * na;
        float* q = Q + b * nq;
        for (long e = t; e < na; e += 1024) r[e] = a[e];
        for (long e = t; e < nq; e += 1024) {
            int i = (int)(e / rows), j = (int)(e - (long)i * rows);
            q[e] = (i == j) ? 1.0f : 0.0f;
        }
        __syncthreads();
        qr_factor<1024, VARIANT>(r, cols, q, rows, rows, cols, t, cs, sync);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <int N, int VARIANT, bool FUSED>
static cudaError_t launch_qr_tpm(const float* A, float* Q, float* R, long batch, cudaStream_t st) {
    constexpr int WARPS = (N == 8) ? 4 : 8;
    constexpr int kSlot = N * N + 4;
    const size_t smem = (size_t)WARPS * 2 * 32 * kSlot * 4 + WARPS * sizeof(uint64_t);
    auto kern = qr_tpm_kernel<N, VARIANT, WARPS, FUSED>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + 31) / 32;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, Q, R, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, int VARIANT>
static cudaError_t launch_qr_cols(const float* A, float* Q, float* R, long batch, cudaStream_t st) {
    constexpr int MPW = 32 / N;
    constexpr int WARPS = (N == 16) ? 16 : 12;
    const size_t smem = (size_t)WARPS * MPW * 2 * (N * N + 4) * 4 + WARPS * sizeof(uint64_t);
    auto kern = qr_cols_kernel<N, VARIANT, WARPS>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + MPW - 1) / MPW;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, Q, R, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, int K, int WARPS, bool FUSED>
static cudaError_t launch_qr_wcols(const float* A, float* Q, float* R, long batch, cudaStream_t st) {
    constexpr int MPW = 32 / (N / K);
    const size_t smem = (size_t)WARPS * MPW * 2 * (N * N + 4) * 4 + WARPS * sizeof(uint64_t);
    // one CTA-wide barrier per wave step (LOCKSTEP) measured +8 % at 32x32 (43.0 M -> 46.7 M fact/s), +1 % at 16x16
    constexpr bool kLock = N == 32;
    auto kern = qr_wcols_kernel<N, K, WARPS, FUSED, kLock>;
    if (N >= 16 && option(HLSD_OPT_LOCKSTEP) == 1) kern = qr_wcols_kernel<N, K, WARPS, FUSED, (N >= 16 && !kLock)>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long num_slabs = (batch + MPW - 1) / MPW;
    long grid = std::min<long>((num_slabs + WARPS - 1) / WARPS, (long)sm_count());
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(A, Q, R, batch);
    count_launch();
    return cudaGetLastError();
}

template <int N, int RT, bool FUSED, bool UNI = false>
static cudaError_t launch_qr_qreg(const float* A, float* Q, float* R, long batch, cudaStream_t st) {
    const size_t smem = ((size_t)N * (N | 1) + 2 * (N / 2 + 2 + 8 + 4 * (RT / 32))) * 4;  // slack behind cs: 3 * GY entries
    auto kern = qr_qreg_kernel<N, RT, FUSED, UNI>;
    constexpr int NT = UNI ? N : RT + N;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    long grid = std::min<long>(batch, (long)sm_count() * std::max(1, per_sm));
    kern<<<(unsigned)grid, NT, smem, st>>>(A, Q, R, batch);
    count_launch();
    return cudaGetLastError();
}

static size_t qr_smem_per_matrix(int rows, int cols) {
    return ((size_t)rows * (cols | 1) + (size_t)rows * (rows | 1) + 2 * (rows / 2 + 2)) * 4;
}

template <int G, int VARIANT>
static cudaError_t launch_qr_wave(const float* A, float* Q, float* R, int rows, int cols, long batch, cudaStream_t st) {
    constexpr int kBlock = G > 256 ? G : 256;
    constexpr int kGroups = kBlock / G;
    const size_t smem = qr_smem_per_matrix(rows, cols) * kGroups;
    auto kern = qr_wave_kernel<G, VARIANT>;
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(227 * 1024) / (smem + 1024)));
    long grid = std::min<long>((batch + kGroups - 1) / kGroups, (long)sm_count() * ctas_per_sm);
    kern<<<(unsigned)grid, kBlock, smem, st>>>(A, Q, R, rows, cols, batch);
    count_launch();
    return cudaGetLastError();
}

template <int VARIANT, bool FUSED>
static cudaError_t qr_dispatch(const float* A, float* Q, float* R, int rows, int cols, long batch, int path,
                               cudaStream_t st) {
    const bool aligned = (((uintptr_t)A | (uintptr_t)Q | (uintptr_t)R) & 15) == 0;
    if (path == HLSD_PATH_AUTO || path == HLSD_PATH_TPM) {
        if (aligned && rows == 4 && cols == 4) return launch_qr_tpm<4, VARIANT, FUSED>(A, Q, R, batch, st);
        if (aligned && rows == 8 && cols == 8) {
            // two lanes per matrix, wavefront: 3.68 G fact/s against 2.95 G for one thread per matrix (batch 4M)
            if constexpr (VARIANT == 0) return launch_qr_wcols<8, 4, 16, FUSED>(A, Q, R, batch, st);
            else return launch_qr_tpm<8, VARIANT, false>(A, Q, R, batch, st);
        }
        // measured (fact/s, batch 1M / 256k): 16x16 K = 1 / 2 / 4 columns per lane 325 M / 463 M / 390 M against 157 M for the
        // sequential lane-per-column kernel; 32x32 K = 1 / 2 39 M / 42 M against 29 M for the shared-memory wavefront
        if (aligned && rows == 16 && cols == 16) {
            if constexpr (VARIANT == 0) return launch_qr_wcols<16, 2, 16, FUSED>(A, Q, R, batch, st);
            else return launch_qr_cols<16, VARIANT>(A, Q, R, batch, st);
        }
        if (aligned && rows == 32 && cols == 32) {
            if constexpr (VARIANT == 0) return launch_qr_wcols<32, 2, 12, FUSED>(A, Q, R, batch, st);
            else if (path == HLSD_PATH_TPM) return launch_qr_cols<32, VARIANT>(A, Q, R, batch, st);
        }
        if (path == HLSD_PATH_TPM) return cudaErrorInvalidValue;
    }
    if (VARIANT == 0 && aligned && rows == cols && path == HLSD_PATH_AUTO) {
        // measured (round 2): one set of threads for R and Q (UNI) 6.22 M / 683 k fact/s at 64 / 128, separate R and Q
        // threads 5.20 M / 541 k
        if (rows == 64 && option(HLSD_OPT_QR_THREADS) != 1) return launch_qr_qreg<64, 64, FUSED, true>(A, Q, R, batch, st);
        if (rows == 128 && option(HLSD_OPT_QR_THREADS) != 1) return launch_qr_qreg<128, 128, FUSED, true>(A, Q, R, batch, st);
        if (rows == 64) return launch_qr_qreg<64, 64, FUSED>(A, Q, R, batch, st);  // (128 R threads, 3 CTAs per SM: 4.0 M vs 5.2 M fact/s)
        // (256 R threads would halve the R side of a step, but 384 threads cap the kernel at 168 registers and the
        // 128-register Q rows spill: 217 k instead of 545 k fact/s)
        if (rows == 128) return launch_qr_qreg<128, 128, FUSED>(A, Q, R, batch, st);
    }
    const size_t per = qr_smem_per_matrix(rows, cols);
    const size_t cap = 227 * 1024 - 1024;
    if ((path == HLSD_PATH_AUTO && per <= cap) || path == HLSD_PATH_GROUP) {
        if (per > cap) return cudaErrorInvalidValue;
        if (rows <= 32 && per * 8 <= cap) return launch_qr_wave<32, VARIANT>(A, Q, R, rows, cols, batch, st);
        if (rows <= 64 && per * 2 <= cap) return launch_qr_wave<128, VARIANT>(A, Q, R, rows, cols, batch, st);
        if (VARIANT == 0 && rows >= 96) return launch_qr_wave<1024, VARIANT>(A, Q, R, rows, cols, batch, st);
        return launch_qr_wave<256, VARIANT>(A, Q, R, rows, cols, batch, st);
    }
    long grid = std::min<long>(batch, (long)sm_count() * 2);
    const size_t smem = (size_t)(2 * (rows / 2 + 2)) * 4;
    qr_global_kernel<VARIANT><<<(unsigned)grid, 1024, smem, st>>>(A, Q, R, rows, cols, batch);
    count_launch();
    return cudaGetLastError();
}

cudaError_t qr_exact(const float* A, float* Q, float* R, int rows, int cols, long batch, int variant, int path, int fused,
                     cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    ProfScope ps(HLSD_PROF_EXACT, st);
    if (variant == 1) return qr_dispatch<1, false>(A, Q, R, rows, cols, batch, path, st);  // qr_v1.x: exact only
    return fused ? qr_dispatch<0, true>(A, Q, R, rows, cols, batch, path, st)
                 : qr_dispatch<0, false>(A, Q, R, rows, cols, batch, path, st);
}

}  // namespace hlsd
