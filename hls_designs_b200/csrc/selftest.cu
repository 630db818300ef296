// Device self-checks exported through the C ABI (used by the GPU tests, not by the data path).
#include "common.cuh"
#include "kernels.h"

namespace hlsd {

// shared-divisor division (common.cuh) against the compiler's IEEE division, bit for bit
__global__ void selftest_division_kernel(const float* __restrict__ num, const float* __restrict__ den, long count,
                                         unsigned long long* __restrict__ out) {
    unsigned long long bad = 0, fast = 0, ungated = 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const float a = num[i], b = den[i];
        const SharedDivisor d = make_divisor(b);
        const float q = xdiv(a, d), ref = __fdiv_rn(a, b);
        bad += __float_as_uint(q) != __float_as_uint(ref);
        fast += d.ok && div_in_range(a);
        // informational: what the three-operation sequence alone would have produced for pairs inside the exponent
        // range, all-ones divisor mantissas included (shows whether this GPU's MUFU.RCP triggers the false tie)
        if (div_in_range(a) && div_in_range(b)) ungated += __float_as_uint(fast_quotient(a, d)) != __float_as_uint(ref);
    }
    if (bad) atomicAdd(out, bad);
    if (fast) atomicAdd(out + 1, fast);
    if (ungated) atomicAdd(out + 2, ungated);
}

cudaError_t selftest_division(const float* num, const float* den, long count, long long* out2 /* 3 counters */, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 3 * sizeof(long long), st);
    if (e != cudaSuccess) return e;
    selftest_division_kernel<<<sm_count() * 8, 256, 0, st>>>(num, den, count, reinterpret_cast<unsigned long long*>(out2));
    count_launch();
    return cudaGetLastError();
}

}  // namespace hlsd
