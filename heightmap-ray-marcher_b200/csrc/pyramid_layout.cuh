// Memory layout of the 16-bit height pyramid (what the march of main/hmap.cpp:1013-1014 fetches from here).
//
// A 32-byte sector — the unit every cache level and HBM move — holds 16 u16 texels.  Row-major, that is a 16 x 1
// strip of the map: a ray (or the 32 rays of an 8x4 screen tile) moving across rows touches a new sector per texel.
// The tiled layouts make a sector a 4 x 4 block of texels instead, so that motion in any direction crosses a sector
// edge every 4 texels (mean crossings per texel over all headings: 0.68 row-major, 0.32 tiled).
//
//   HMRM_LAYOUT_ROWMAJOR  idx = y * w + x                                                         (round 1)
//   HMRM_LAYOUT_TILE4     4x4-texel sectors, sectors in row-major order:
//                         idx = ((y >> 2) * spr + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3),  spr = ceil(w / 4)
//   HMRM_LAYOUT_ZORDER    4x4-texel sectors in Z-order (Morton) inside 64x64-texel (8 KiB) blocks, blocks row-major:
//                         sector coordinates (x >> 2, y >> 2), their low 4 bits interleaved y3 x3 y2 x2 y1 x1 y0 x0
//
// One level is described by a uint2 (element offset of the level, layout-specific pitch term), so that the kernels
// read both from the constant bank with one indexed load:
//   ROWMAJOR  pitch = w
//   TILE4     pitch = 16 * spr - 16     (idx = x + 12 (x >> 2) + 4 y + pitch (y >> 2))
//   ZORDER    pitch = blocks per row    (64x64 blocks)
// K1 writes with the same functions K2 reads with.
#ifndef HMRM_PYRAMID_LAYOUT_CUH
#define HMRM_PYRAMID_LAYOUT_CUH

#include <stdint.h>

namespace hmrm {

enum { kLayoutRowMajor = 0, kLayoutTile4 = 1, kLayoutZOrder = 2 };

// host + device: number of u16 elements a w x h level occupies, and its pitch term
__host__ __device__ inline size_t pyr_level_elems(int layout, int w, int h) {
	if (layout == kLayoutTile4) return (size_t)((w + 3) / 4) * (size_t)((h + 3) / 4) * 16;
	if (layout == kLayoutZOrder) return (size_t)((w + 63) / 64) * (size_t)((h + 63) / 64) * 4096;
	return (size_t)w * (size_t)h;
}
__host__ __device__ inline unsigned pyr_level_pitch(int layout, int w) {
	if (layout == kLayoutTile4) return 16u * (unsigned)((w + 3) / 4) - 16u;
	if (layout == kLayoutZOrder) return (unsigned)((w + 63) / 64);
	return (unsigned)w;
}

// spread the low 4 bits of v to the even bit positions: abcd -> 0a0b0c0d
__host__ __device__ inline unsigned pyr_spread4(unsigned v) {
	v = (v | (v << 2)) & 0x33u;
	return (v | (v << 1)) & 0x55u;
}

// element index of texel (x, y) inside its level
template <int kLayout>
__host__ __device__ __forceinline__ unsigned pyr_index(unsigned x, unsigned y, unsigned pitch) {
	if (kLayout == kLayoutTile4) return x + 12u * (x >> 2) + 4u * y + pitch * (y >> 2);
	if (kLayout == kLayoutZOrder) {
		const unsigned sx = x >> 2, sy = y >> 2;
		const unsigned block = (sy >> 4) * pitch + (sx >> 4);
		const unsigned sector = pyr_spread4(sx & 15u) | (pyr_spread4(sy & 15u) << 1);
		return block * 4096u + sector * 16u + (y & 3u) * 4u + (x & 3u);
	}
	return y * pitch + x;
}

__host__ __device__ __forceinline__ unsigned pyr_index_rt(int layout, unsigned x, unsigned y, unsigned pitch) {
	if (layout == kLayoutTile4) return pyr_index<kLayoutTile4>(x, y, pitch);
	if (layout == kLayoutZOrder) return pyr_index<kLayoutZOrder>(x, y, pitch);
	return pyr_index<kLayoutRowMajor>(x, y, pitch);
}

} // namespace hmrm

#endif
