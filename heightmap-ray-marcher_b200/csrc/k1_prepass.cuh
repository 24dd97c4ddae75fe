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
                                                  unsigned long long *__restrict__ max_surf_bits,
                                                  unsigned long long *__restrict__ min_surf_bits) {
	unsigned long long local_max = 0ULL, local_min = ~0ULL;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const uint8_t *px = rgb8 + 3 * p;
		const double h = height_of(q, px[0], px[1], px[2]);
		if (heights) heights[p] = h;
		const double s = fadd(h, q.min_height);
		if (surf) surf[p] = s;
		const unsigned long long ob = ordered_bits(s);
		if (ob > local_max) local_max = ob;
		if (ob < local_min) local_min = ob;
	}
	if (max_surf_bits) {
		for (int off = 16; off > 0; off >>= 1) {
			const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, local_max, off);
			if (o > local_max) local_max = o;
		}
		if ((threadIdx.x & 31) == 0 && local_max != 0ULL) atomicMax(max_surf_bits, local_max);
	}
	if (min_surf_bits) {
		for (int off = 16; off > 0; off >>= 1) {
			const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, local_min, off);
			if (o < local_min) local_min = o;
		}
		if ((threadIdx.x & 31) == 0 && local_min != ~0ULL) atomicMin(min_surf_bits, local_min);
	}
}

// ---- conservative fixed-point view of the surface (used by the skip traversal) -------------
//
// Zq(v) = low 32 bits of fma(v, zq_scale, zq_offset) where zq_offset = 1.5*2^52 + c: the classic
// magic-number conversion, one DFMA instead of a quarter-rate F2I.  Zq is monotone non-decreasing
// in v (zq_scale > 0, rounding is monotone), and K1 and K2 evaluate the very same instruction, so
//     Zq(z) > Zq(surf)  =>  z > surf   (the march's test `z < surf` is false)
//     Zq(z) < Zq(surf)  =>  z < surf   (hit)
// with no error analysis; only Zq(z) == Zq(surf) needs the FP64 surf value.
#define HMRM_MAGIC 6755399441055744.0   /* 1.5 * 2^52 */

// low word of (magic + n) is n for |n| < 2^31; the high word tells whether that held
__device__ __forceinline__ bool magic_decode(double t, int &v) {
	const int hi = __double2hiint(t);
	v = __double2loint(t);
	return (hi == 0x43380000 && v >= 0) || (hi == 0x4337FFFF && v < 0);
}

__device__ __forceinline__ int zq_of(double z, double zq_scale, double zq_offset) {
	int v;
	const double t = __fma_rn(z, zq_scale, zq_offset);
	if (magic_decode(t, v)) return v;
	return (t > HMRM_MAGIC) ? INT_MAX : INT_MIN;   // far above / far below every surface value (NaN: below)
}

// q0[p] = Zq(surf[p]); flags an error if a value does not fit 16 bits (cannot happen with the host's scale)
__global__ void __launch_bounds__(256) k1_quantise(const double *__restrict__ surf, long long n, double zq_scale,
                                                   double zq_offset, uint16_t *__restrict__ q0,
                                                   unsigned int *__restrict__ error_flag) {
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const int q = zq_of(surf[p], zq_scale, zq_offset);
		if (q < 0 || q > 65535) {
			atomicOr(error_flag, 1u);
			q0[p] = 65535;
		}
		else q0[p] = (uint16_t)q;
	}
}

// dst[y][x] = max of the 2x2 block of src (clipped at the edges): one mip level
__global__ void __launch_bounds__(256) k1_mip_reduce(const uint16_t *__restrict__ src, int sw, int sh,
                                                     uint16_t *__restrict__ dst, int dw, int dh) {
	const long long n = (long long)dw * dh;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const int x = (int)(p % dw), y = (int)(p / dw);
		const int x0 = 2 * x, y0 = 2 * y;
		const int x1 = (x0 + 1 < sw) ? x0 + 1 : x0, y1 = (y0 + 1 < sh) ? y0 + 1 : y0;
		const uint16_t a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1];
		const uint16_t c = src[(size_t)y1 * sw + x0], d = src[(size_t)y1 * sw + x1];
		const uint16_t ab = a > b ? a : b, cd = c > d ? c : d;
		dst[p] = ab > cd ? ab : cd;
	}
}

// dst[y][x] = max of src over the 3x3 neighbourhood (clipped): a sample that clears dst clears the block it is in
// AND the eight blocks around it, so a jump may run on into the neighbouring blocks instead of stopping at an edge.
__global__ void __launch_bounds__(256) k1_mip_dilate(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst,
                                                     int w, int h) {
	const long long n = (long long)w * h;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
		const int x = (int)(p % w), y = (int)(p / w);
		uint16_t m = 0;
		for (int dy = -1; dy <= 1; ++dy) {
			const int yy = y + dy;
			if (yy < 0 || yy >= h) continue;
			for (int dx = -1; dx <= 1; ++dx) {
				const int xx = x + dx;
				if (xx < 0 || xx >= w) continue;
				const uint16_t v = src[(size_t)yy * w + xx];
				m = v > m ? v : m;
			}
		}
		dst[p] = m;
	}
}

} // namespace hmrm

#endif
