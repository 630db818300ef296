// Shared device/host helpers for the batched decomposition kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace hlsd {

// compile-time loop: f(std::integral_constant<int, I>) for I in [Begin, End); every index inside the
// body is a constant expression, so register arrays indexed by it never fall back to local memory
template <int Begin, int End, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (Begin < End) {
        f(std::integral_constant<int, Begin>{});
        static_for<Begin + 1, End>(f);
    }
}

// ----------------------------------------------------------------------------------------------
// Exact arithmetic.  The reference's C-simulation rounds every float operation separately
// (g++ on baseline x86-64: no FMA; e.g. `mid_L[j][i] - mid_L[j][id] * mid_L[i][id]`,
// Cholesky/cholesky_v1.3/model4x4/chol.cpp:59).  The intrinsics below are never contracted by
// nvcc into FFMA and use IEEE division / square root, so a kernel that performs the reference's
// operations in the reference's per-element order reproduces its results bit for bit.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
// IEEE division of several numerators by one divisor.  The reciprocal and its Newton step are computed once;
// each quotient then costs three operations: q0 = a*r, rem = a - b*q0 (exact, fused), q = q0 + rem*r.  The
// result is the correctly rounded quotient whenever the refined reciprocal r equals RN(1/b) (Markstein), which
// the Newton step guarantees from any MUFU.RCP value within 2 ulp - EXCEPT for divisors whose mantissa is all
// ones, where r can stay one ulp low and a power-of-two numerator then lands on a false tie
// (0x1p23 / 0x1.fffffep23 -> 0x1p-1 instead of 0x1.000002p-1; round-1 advisor finding, reproduced on the CPU
// with glibc fmaf: 65 694 mismatches in 96 M hard cases with those divisors, 0 in 154 M without them, and a
// second residual correction does not help because the residual is exact and r is what is off).  Such
// divisors, zeros, subnormals, infinities, NaNs and extreme exponents (`div_in_range`, a conservative
// sub-range of what nvcc's own fast path accepts) fall back to __fdiv_rn.  tests/test_gpu_division.py checks
// the quotients bit for bit against __fdiv_rn on the device, including all-ones and near-all-ones divisors.
struct SharedDivisor {
    float b, r;
    bool ok;
};
__device__ __forceinline__ bool div_in_range(float x) {
    const float ax = fabsf(x);
    return ax >= 0x1p-60f && ax <= 0x1p60f;
}
__device__ __forceinline__ SharedDivisor make_divisor(float b) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b));
    const float e = __fmaf_rn(-b, r0, 1.0f);
    const bool all_ones = (__float_as_uint(b) & 0x7fffffu) == 0x7fffffu;
    return SharedDivisor{b, __fmaf_rn(r0, e, r0), div_in_range(b) && !all_ones};
}
__device__ __forceinline__ float fast_quotient(float a, const SharedDivisor& d) {
    const float q0 = __fmul_rn(a, d.r);
    const float rem = __fmaf_rn(-d.b, q0, a);
    return __fmaf_rn(d.r, rem, q0);
}
// one numerator (the checked form)
__device__ __forceinline__ float xdiv(float a, const SharedDivisor& d) {
    return (d.ok && div_in_range(a)) ? fast_quotient(a, d) : __fdiv_rn(a, d.b);
}

// a - b*c, two roundings
__device__ __forceinline__ float xmsub(float a, float b, float c) { return __fsub_rn(a, __fmul_rn(b, c)); }

// the same with the contraction HLSD_FLAG_FUSED allows (FUSED: one rounding)
template <bool FUSED>
__device__ __forceinline__ float msub(float a, float b, float c) {
    if constexpr (FUSED) return __fmaf_rn(-b, c, a);
    else return xmsub(a, b, c);
}

// ----------------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100a FMUL2 / FFMA2): one instruction works on two floats held in an aligned register pair,
// each lane rounded like the scalar instruction.  The exact kernels are bound by instruction issue and fetch, not by
// the FP32 pipe, so the separately rounded a - b*c costs half the instructions this way:  m = b*c (FMUL2, rounded),
// then a - m as fma(m, -1, a): the product m * -1 is exact, so the single rounding of the sum is the rounding of the
// subtraction.  (mul.f32x2 followed directly by sub.f32x2 is NOT safe: ptxas contracts that pair into FFMA2 even
// with --fmad=false; it cannot contract a product that feeds the multiplicand of an fma.)
// ----------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 xmul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// (a.lo - b.lo*c.lo, a.hi - b.hi*c.hi), two roundings per lane; FUSED: one (nb = -b must be passed as well)
template <bool FUSED>
__device__ __forceinline__ f32x2 msub2(f32x2 a, f32x2 b, f32x2 nb, f32x2 c) {
    if constexpr (FUSED) return fma2(nb, c, a);
    else return fma2(xmul2(b, c), 0xBF800000BF800000ULL, a);
}
// the same on four scalars: (x0, x1) -= b * (c0, c1)
template <bool FUSED>
__device__ __forceinline__ void msub_pair(float& x0, float& x1, float b, float c0, float c1) {
    const f32x2 r = msub2<FUSED>(pack2(x0, x1), pack2(b, b), pack2(-b, -b), pack2(c0, c1));
    unpack2(r, x0, x1);
}

// qrf_mm of the reference (QR/qr_v2.1/model4x4/qr.cpp:52-60): a' = op2*s + op1*c ; b' = op2*c - op1*s
__device__ __forceinline__ void givens_apply(float c, float s, float& op1, float& op2) {
    float a = xadd(xmul(op2, s), xmul(op1, c));
    float b = xsub(xmul(op2, c), xmul(op1, s));
    op1 = a;
    op2 = b;
}
template <bool FUSED>
__device__ __forceinline__ void givens_apply_t(float c, float s, float& op1, float& op2) {
    if constexpr (FUSED) {
        const float a = __fmaf_rn(op2, s, __fmul_rn(op1, c));
        const float b = __fmaf_rn(op2, c, -__fmul_rn(op1, s));
        op1 = a;
        op2 = b;
    } else {
        givens_apply(c, s, op1, op2);
    }
}

// ----------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copies (TMA engine, SASS: UBLKCP).  Sizes and both addresses must be
// multiples of 16 bytes.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared::cta, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared::cta -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of all but the newest `N` groups have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// make generic-proxy writes to shared memory visible to the async proxy (before bulk_s2g)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Cooperative, fully coalesced copy of a slab of `valid` (<= 32) small matrices between global memory
// (dense, F4 float4 per matrix) and shared memory slots padded to F4 + 1 float4 (the padding that makes the
// per-thread 16-byte accesses of the thread-per-matrix kernels conflict-free).  For 64- and 256-byte
// matrices this beats one bulk copy per matrix: a warp instruction moves 512 contiguous bytes.
template <int F4>
__device__ __forceinline__ void slab_g2s(float4* slab, const float4* __restrict__ g, int valid, int lane) {
#pragma unroll
    for (int it = 0; it < F4; it++) {
        const int idx = it * 32 + lane, m = idx / F4, q = idx % F4;
        if (m < valid) slab[m * (F4 + 1) + q] = __ldg(g + idx);
    }
}
template <int F4>
__device__ __forceinline__ void slab_s2g(float4* __restrict__ g, const float4* slab, int valid, int lane) {
#pragma unroll
    for (int it = 0; it < F4; it++) {
        const int idx = it * 32 + lane, m = idx / F4, q = idx % F4;
        if (m < valid) __stcs(g + idx, slab[m * (F4 + 1) + q]);
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Group of G threads (G a power of two, 32 <= G <= 1024, groups aligned inside the CTA) that
// works on one matrix.  A warp-sized group synchronises with __syncwarp, larger ones with a
// named barrier (ids 1..15; id 0 is __syncthreads).
template <int G>
__device__ __forceinline__ void group_sync(int group_in_cta) {
    if constexpr (G == 32) __syncwarp();
    else named_bar_sync(1 + group_in_cta, G);
}

}  // namespace hlsd
