// Microbenchmark: un-fused multiply-subtract a - b*c (two roundings) on 32 registers per thread, scalar
// (FMUL + FADD) against packed (FMUL2 + FFMA2 with a (-1,-1) multiplier, which ptxas cannot contract).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* __restrict__ out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] = in[(t + j * 7) & 1023];
    float l = in[t & 1023] * 1e-3f;
    if (MODE == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = __fsub_rn(a[j], __fmul_rn(l, a[(j + 1) & 31]));
        }
    } else {
        u64 p[16];
#pragma unroll
        for (int j = 0; j < 16; j++) p[j] = pack(a[2 * j], a[2 * j + 1]);
        const u64 ll = pack(l, l), m1 = pack(-1.0f, -1.0f);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 16; j++) p[j] = fma2(mul2(ll, p[(j + 1) & 15]), m1, p[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) unpack(p[j], a[2 * j], a[2 * j + 1]);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) s += a[j];
    out[t] = s;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096);
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    float h[1024];
    for (int i = 0; i < 1024; i++) h[i] = 1.0f + i * 1e-3f;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    const int iters = 4096;
    for (int mode = 0; mode < 2; mode++) {
        cudaEvent_t a, b;
        cudaEventCreate(&a), cudaEventCreate(&b);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            if (mode == 0) k<0><<<148 * 8, 256>>>(in, out, iters);
            else k<1><<<148 * 8, 256>>>(in, out, iters);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
        }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double pairs = 148.0 * 8 * 256 * 32 * iters;
        printf("%s: %.3f ms, %.2f T (mul,sub) pairs/s\n", mode == 0 ? "scalar FMUL+FADD" : "packed FMUL2+FFMA2", ms, pairs / ms * 1e-9);
    }
    float r[2];
    cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost);
    printf("%g %g %s\n", r[0], r[1], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
