// K1 — one-shot height prepass (replaces UpdateHeightmap, main/hmap.cpp:171-191).
//
//   v      = clamp((lum_r*R + lum_g*G) + lum_b*B, 0, 255)
//   height = (v / 255) * (max_height - min_height) + min_height            (:187-188)
//   surf   = height + min_height     <- what the march compares against     (:1016, hmap_c0.z = min_height)
//
// The reference keeps `height` and adds min_height again at every march step;
// K1 stores `surf` (same two roundings, done once per cell) plus the global
// maximum of surf, which bounds every terrain sample from above.
#ifndef HMRM_K1_PREPASS_CUH
#define HMRM_K1_PREPASS_CUH

#include "device_math.cuh"

namespace hmrm {

struct PrepassParams {
	double lum_r, lum_g, lum_b;
	double min_height;
	double span;          // fl(max_height - min_height)
};

__device__ __forceinline__ double height_of(const PrepassParams &q, uint32_t r, uint32_t g, uint32_t b) {
	double v = fadd(fadd(fmul(q.lum_r, (double)r), fmul(q.lum_g, (double)g)), fmul(q.lum_b, (double)b));
	if (v < 0.0) v = 0.0;
	else if (v > 255.0) v = 255.0;
	return fadd(fmul(fdiv(v, 255.0), q.span), q.min_height);
}

// order-preserving map double -> uint64 so that atomicMax works on FP64 values
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
	const unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ inline double from_ordered_bits(unsigned long long o) {
	const unsigned long long b = (o & 0x8000000000000000ULL) ? (o & 0x7FFFFFFFFFFFFFFFULL) : ~o;
	double v;
#if defined(__CUDA_ARCH__)
	v = __longlong_as_double((long long)b);
#else
	__builtin_memcpy(&v, &b, sizeof v);
#endif
	return v;
}

// rgb8: [n][3]; surf: [n] (may be NULL); heights: [n] (may be NULL, parity checks only)
__global__ void __launch_bounds__(256) k1_prepass(const uint8_t *__restrict__ rgb8, long long n, PrepassParams q,
                                                  double *__restrict__ surf, double *__restrict__ heights,
                                                  unsigned long long *__restrict__ max_surf_bits) {
	unsigned long long local_max = 0ULL;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const uint8_t *px = rgb8 + 3 * p;
		const double h = height_of(q, px[0], px[1], px[2]);
		if (heights) heights[p] = h;
		const double s = fadd(h, q.min_height);
		if (surf) surf[p] = s;
		const unsigned long long ob = ordered_bits(s);
		if (ob > local_max) local_max = ob;
	}
	if (max_surf_bits) {
		for (int off = 16; off > 0; off >>= 1) {
			const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, local_max, off);
			if (o > local_max) local_max = o;
		}
		if ((threadIdx.x & 31) == 0 && local_max != 0ULL) atomicMax(max_surf_bits, local_max);
	}
}

} // namespace hmrm

#endif
