// K2 (brute traversal) — fused ray generation + AABB entry + fixed-step march +
// colormap fetch, one height fetch per reference step: main/hmap.cpp:978-1058
// restated for a GPU.  This is the baseline kernel of the exact mode: it issues
// exactly the fetches the reference issues, so its step count is the
// reference-equivalent step count S and its memory behaviour is the un-skipped
// fetch roofline.  The production traversal (k2_render_skip.cuh) must produce
// the same framebuffer and the same per-pixel first-hit step index.
//
// Scheduling: 8x4-pixel screen tiles, one warp per tile, persistent CTAs that
// pull tile ids from a global atomic counter (march lengths are bimodal, sky
// tiles cost ~nothing, terrain tiles cost thousands of steps).
#ifndef HMRM_K2_RENDER_BRUTE_CUH
#define HMRM_K2_RENDER_BRUTE_CUH

#include "ray_setup.cuh"

namespace hmrm {

// One warp-wide grab from the tile queue.
__device__ __forceinline__ unsigned next_tile(unsigned int *counter) {
	unsigned t = 0;
	if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, 1u);
	return __shfl_sync(0xFFFFFFFFu, t, 0);
}

// tile -> (row, column) of the launch's tile grid.  The float estimate of the quotient never exceeds it (reciprocal
// and conversions round down; tile < 2^26 so the conversion is exact below 2^24 and off by < 4 above, where
// tiles_x >= 1024), and is short by at most a couple of units: fixed up with compares instead of an integer divide.
__device__ __forceinline__ void tile_row_col(const RenderParams &P, unsigned tile, int &row, int &col) {
	unsigned q = __float2uint_rz(__uint2float_rz(tile) * P.inv_tiles_x);
	unsigned r = tile - q * (unsigned)P.tiles_x;
	while (r >= (unsigned)P.tiles_x) {
		r -= (unsigned)P.tiles_x;
		q += 1u;
	}
	row = (int)q;
	col = (int)r;
}

// Does pixel (px,py) belong to this frame?  (cycle interleave main/hmap.cpp:979-981, row band)
__device__ __forceinline__ bool pixel_selected(const RenderParams &P, int px, int py) {
	if (px >= P.W || py >= P.row_end) return false;
	if (P.period > 1) {
		const long long p = (long long)py * P.W + px;
		if (p < P.cycle || ((p - P.cycle) % P.period) != 0) return false;
	}
	return true;
}

struct PixelTally {
	unsigned long long steps, fetches;
	unsigned box_hit, surf_hit, cut_off;
	unsigned dbg[8];
};

template <bool kStats>
__device__ __forceinline__ void commit_tally(const RenderParams &P, bool active, const PixelTally &t) {
	if (!kStats) return;
	unsigned long long rays = active ? 1ULL : 0ULL, box = t.box_hit, surf = t.surf_hit;
	unsigned long long steps = t.steps, fetches = t.fetches, mx = t.steps;
	unsigned cut = t.cut_off;
	for (int off = 16; off > 0; off >>= 1) {
		rays += __shfl_xor_sync(0xFFFFFFFFu, rays, off);
		box += __shfl_xor_sync(0xFFFFFFFFu, box, off);
		surf += __shfl_xor_sync(0xFFFFFFFFu, surf, off);
		steps += __shfl_xor_sync(0xFFFFFFFFu, steps, off);
		fetches += __shfl_xor_sync(0xFFFFFFFFu, fetches, off);
		const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, mx, off);
		mx = o > mx ? o : mx;
		cut |= __shfl_xor_sync(0xFFFFFFFFu, cut, off);
	}
	for (int i = 0; i < 8; ++i) {
		unsigned long long v = t.dbg[i];
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
		if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.stats->dbg[i], v);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&P.stats->rays, rays);
		atomicAdd(&P.stats->box_hits, box);
		atomicAdd(&P.stats->surf_hits, surf);
		atomicAdd(&P.stats->steps, steps);
		atomicAdd(&P.stats->fetches, fetches);
		atomicMax(&P.stats->max_steps, mx);
		if (cut) atomicOr(&P.stats->status, 4u);   // HMRM_ERR_NONTERMINATING
	}
}

template <bool kStats>
__global__ void __launch_bounds__(256) k2_render_brute(const __grid_constant__ RenderParams P) {
	const int lane = threadIdx.x & 31;
	const unsigned n_tiles = (unsigned)(P.tiles_x * P.tiles_y);

	for (;;) {
		const unsigned tile = next_tile(P.tile_counter);
		if (tile >= n_tiles) break;
		int ty, tx;
		tile_row_col(P, tile, ty, tx);
		const int px = tx * 8 + (lane & 7);
		const int py = P.row_begin + (P.tile_y_first + ty * P.tile_y_step) * 4 + (lane >> 3);
		const bool active = pixel_selected(P, px, py);

		PixelTally tally = {0ULL, 0ULL, 0u, 0u, 0u, {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}};
		uint32_t rgba = 0u;
		if (active) {
			const Ray ray = generate_ray(P, px, py);
			bool real_hit = false;
			int first_hit = -1;
			double x, y, z, lo = 0.0;
			const bool entered = box_entry(P, ray, x, y, z, lo);
			if (kStats && P.ray_dump) dump_ray(P, px, py, ray, entered, lo, x, y, z);
			if (entered) {
				tally.box_hit = 1u;
				first_hit = -2;
				x = fadd(x, fmul(P.nudge, ray.dx));               // main/hmap.cpp:998
				y = fadd(y, fmul(P.nudge, ray.dy));
				z = fadd(z, fmul(P.nudge, ray.dz));
				const double sx = fmul(P.step_dist, ray.dx);       // loop-invariant part of :1037
				const double sy = fmul(P.step_dist, ray.dy);
				const double sz = fmul(P.step_dist, ray.dz);
				unsigned long long k = 0ULL;
				for (;;) {
					const int gx = trunc_cell(fdiv(x, P.gw));      // :1001-1002 (hmap_c0.x == 0)
					const int gy = trunc_cell(fdiv(-y, P.gw));     // :1003-1004 (hmap_c0.y == 0)
					if (gx < 0 || gy < 0 || gx >= P.map_w || gy >= P.map_h) break;
					const size_t cell = (size_t)gx + (size_t)gy * (size_t)P.map_w;
					const double s = __ldg(P.surf + HMRM_CHECKED(P, cell, (size_t)P.map_w * (size_t)P.map_h));
					k += 1ULL;
					if (z < s) {                                    // :1016
						rgba = hit_colour(P, __ldg(P.color + HMRM_CHECKED(P, cell, (size_t)P.map_w * (size_t)P.map_h)));
						real_hit = true;
						first_hit = (k - 1ULL > 0x7FFFFFFFULL) ? 0x7FFFFFFF : (int)(k - 1ULL);
						break;
					}
					const double nx = fadd(x, sx), ny = fadd(y, sy), nz = fadd(z, sz);   // :1037
					// The reference loops forever when a ray can no longer move in x/y and cannot
					// come down onto the cell it is over; cut such rays off and flag the frame.
					if (nx == x && ny == y && !(nz < z)) {
						tally.cut_off = 1u;
						break;
					}
					x = nx; y = ny; z = nz;
				}
				tally.steps = k;
				tally.fetches = k;
			}
			if (!real_hit) rgba = miss_colour(P, ray.dz);
			else tally.surf_hit = 1u;
			if (P.step_index) P.step_index[(size_t)py * (size_t)P.W + (size_t)px] = first_hit;
		}
		store_pixel(P, px, py, active, rgba);            // SetPixel, main/hmap.cpp:139-154
		commit_tally<kStats>(P, active, tally);
	}
}

} // namespace hmrm

#endif
