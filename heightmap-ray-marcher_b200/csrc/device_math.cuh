// Exactly-rounded FP64 primitives for the bit-exact path.
//
// Each wrapper is ONE IEEE-754 binary64 round-to-nearest-even operation and can
// never be contracted into an FMA (the *_rn intrinsics are contraction barriers),
// so the kernels reproduce the reference's x86-64 SSE2 arithmetic independently
// of compiler flags; the library is nevertheless built with --fmad=false.
#ifndef HMRM_DEVICE_MATH_CUH
#define HMRM_DEVICE_MATH_CUH

#include <cuda_runtime.h>
#include <limits.h>

namespace hmrm {

__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fsqrt(double a) { return __dsqrt_rn(a); }
// 1.0 / a: the correctly rounded reciprocal IS the correctly rounded quotient of 1 and a (one rounding of the same
// real number), and costs fewer instructions than the general divide
__device__ __forceinline__ double frcp(double a) { return __drcp_rn(a); }

// (int)double as the reference's x86-64 build computes it (cvttsd2si): NaN and
// out-of-range give INT_MIN.  CUDA's conversion saturates and maps NaN to 0, so
// NaN is handled explicitly; the saturated values are out of the grid either way.
__device__ __forceinline__ int trunc_cell(double q) {
	return (q == q) ? __double2int_rz(q) : INT_MIN;
}

} // namespace hmrm

#endif
