// Blocked right-looking factorisations for large N.  (placeholder until the panel/trailing kernels land)
#include "common.cuh"
#include "kernels.h"

namespace hlsd {
cudaError_t cholesky_blocked(const float*, float*, int, long, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace hlsd
