// CPU stand-ins for the handful of CUDA builtins skeleton.cuh and the emitted bodies use.
// TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` execute the emitted expression bodies and
// the skeleton's step logic with g++ and compare them with the oracle.  Never used by the product.
#ifndef GFB_CUDA_STUB_HPP
#define GFB_CUDA_STUB_HPP
#define GFB_HOST_HARNESS 1
#include <cmath>
#include <cstddef>
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__
#define __align__(x) __attribute__((aligned(x)))
struct double2 { double x, y; };
template<typename T> inline T __ldg(const T *p) { return *p; }
struct gfb_dim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local gfb_dim3 threadIdx, blockIdx, blockDim;
inline void __syncthreads() {}
using std::fma; using std::fmin; using std::fmax; using std::sqrt; using std::fabs;
using std::exp; using std::log; using std::pow; using std::sin; using std::cos; using std::atan2;
inline void sincos(const double x, double *s, double *c) { *s = std::sin(x); *c = std::cos(x); }
//  cvt.rzi.u32.f64: truncation, saturating; negative values and NaN give 0.
inline unsigned __double2uint_rz(const double x) {
    return !(x > 0.0) ? 0u : (x >= 4294967295.0 ? 4294967295u : static_cast<unsigned> (x));
}
inline unsigned min(const unsigned a, const unsigned b) { return a < b ? a : b; }
#endif
