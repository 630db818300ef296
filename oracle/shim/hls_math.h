// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Xilinx's hls_math.h.  In C-simulation the
// float overloads resolve to the host libm (correctly rounded sqrt, exact fabs).
#ifndef ORACLE_SHIM_HLS_MATH_H
#define ORACLE_SHIM_HLS_MATH_H
#include <cmath>
namespace hls {
static inline float abs(float x) { return std::fabs(x); }
static inline double abs(double x) { return std::fabs(x); }
static inline int abs(int x) { return x < 0 ? -x : x; }
static inline float sqrtf(float x) { return std::sqrt(x); }
static inline float sqrt(float x) { return std::sqrt(x); }
static inline double sqrt(double x) { return std::sqrt(x); }
}  // namespace hls
#endif
