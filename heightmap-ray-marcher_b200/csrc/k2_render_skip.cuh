// K2 (skip traversal) — the production kernel of the exact mode.
//
// Same contract as k2_render_brute.cuh (main/hmap.cpp:978-1058): same framebuffer, same per-pixel
// first-hit sample index, same reference-equivalent step count — but most of the reference's height
// fetches are never issued.  Three facts make that exact rather than approximate:
//
// (1) Closed-form stepping.  The reference advances P <- fl(P + s) per axis (main/hmap.cpp:1037), s =
//     fl(step_dist*dir).  While P stays in one binade (same sign and exponent, ulp U) every such add
//     moves P by the same multiple of U, S = fl(fl(P+s) - P), provided two consecutive increments agree
//     (that also rules out the round-half-even alternation when s is an odd multiple of U/2).  Then
//     P_m = P + m*S exactly, and fma(m, S, P) returns it with a single (exact) rounding.  If the end
//     point leaves the binade the jump is refused and retried with m/2; m = 1 is always a plain add.
// (2) Monotone fixed point.  Cell coordinates and heights are compared through
//     low32(fma(v, scale, 1.5*2^52)) — monotone in v, one DFMA instead of a divide plus a quarter-rate
//     F2I.  The cell index is taken from the fixed-point value only when it is at least one unit away
//     from a cell edge (|fixed - 2^k*fl(x/gw)| < 1); otherwise the reference's divide is evaluated.
//     Heights use the same Zq() in K1 and here, so Zq(z) > Zq(surf) => z > surf with no error term.
// (3) Monotone rays.  Within (1) every axis moves monotonically, so if sample i and sample i+m both lie
//     in one mip block B and both have Zq(z) > max_B Zq(surf), every sample between them lies in B
//     (hence in the grid) and above every surface value in B: none of them can hit or leave the grid.
//     The jump length m is only an estimate; the end point is verified before it is committed.
//
// Scheduling is the same persistent 8x4-tile queue as the brute kernel.
#ifndef HMRM_K2_RENDER_SKIP_CUH
#define HMRM_K2_RENDER_SKIP_CUH

#include "k1_prepass.cuh"
#include "k2_render_brute.cuh"

namespace hmrm {

#define HMRM_JUMP_CAP (1 << 20)

__device__ __forceinline__ int binade_tag(double v) { return __double2hiint(v) >> 20; }   // sign + exponent

// exact double of a small non-negative int without the quarter-rate I2F: (2^52 + m) - 2^52
__device__ __forceinline__ double small_int_to_double(int m) {
	return fsub(__hiloint2double(0x43300000, m), 4503599627370496.0);
}

struct AxisState {
	double p;     // position component
	double s;     // fl(step_dist * dir) component: the reference's per-step addend
	double S;     // increment per step while p stays in binade `tag`
	int tag;      // binade for which S is valid, or INT_MIN
};

// (1): find the constant per-step increment of this axis in p's binade, or mark it unknown.
__device__ __forceinline__ void axis_refresh(AxisState &a) {
	const double t1 = fadd(a.p, a.s);
	const double t2 = fadd(t1, a.s);
	const double d1 = fsub(t1, a.p);
	const double d2 = fsub(t2, t1);
	const int e0 = binade_tag(a.p);
	if (e0 == binade_tag(t1) && e0 == binade_tag(t2) && d1 == d2) {
		a.S = d1;
		a.tag = e0;
	}
	else a.tag = INT_MIN;
}

// make sure S is known for p's current binade; false if it cannot be established right now
__device__ __forceinline__ bool axis_ready(AxisState &a) {
	if (a.tag != binade_tag(a.p)) axis_refresh(a);
	return a.tag != INT_MIN;
}

// end point of m steps on this axis (axis_ready held); false if it would leave the binade
__device__ __forceinline__ bool axis_jump(const AxisState &a, double md, double &end) {
	end = __fma_rn(md, a.S, a.p);
	return binade_tag(end) == a.tag;
}

// Cell index along one axis from the fixed-point coordinate t = fma(c, 2^k/gw, magic); `c` is x or -y.
// Returns the reference's (int)(c / grid_width) (main/hmap.cpp:1001-1004), or a value outside [0, dim)
// when the sample is outside the grid on that side.  `v` receives the raw fixed-point value.
__device__ __forceinline__ int axis_cell(double c, double t, int k, double gw, int &v) {
	if (!magic_decode(t, v)) {
		// |c/gw| >= 2^(31-k) >= map dimension, or NaN: outside the grid either way
		const bool above = t > HMRM_MAGIC;
		v = above ? INT_MAX : INT_MIN;
		return above ? INT_MAX : -1;
	}
	const unsigned mask = (1u << k) - 1u;
	if ((((unsigned)v + 1u) & mask) <= 1u) return trunc_cell(fdiv(c, gw));   // within one unit of a cell edge
	if (v < 0) return (v > -(1 << k)) ? 0 : -1;                                // (int) truncates toward zero
	return v >> k;
}

template <bool kStats>
__global__ void __launch_bounds__(256) k2_render_skip(const __grid_constant__ RenderParams P) {
	const int lane = threadIdx.x & 31;
	const unsigned n_tiles = (unsigned)(P.tiles_x * P.tiles_y);
	const int k = P.fx_bits;

	for (;;) {
		const unsigned tile = next_tile(P.tile_counter);
		if (tile >= n_tiles) break;
		const int ty = (int)(tile / (unsigned)P.tiles_x);
		const int tx = (int)(tile - (unsigned)ty * (unsigned)P.tiles_x);
		const int px = tx * 8 + (lane & 7);
		const int py = P.row_begin + ty * 4 + (lane >> 3);
		const bool active = pixel_selected(P, px, py);

		PixelTally tally = {0ULL, 0ULL, 0u, 0u, 0u};
		if (active) {
			const Ray ray = generate_ray(P, px, py);
			uint32_t rgba = 0u;
			bool real_hit = false;
			int first_hit = -1;
			double ex, ey, ez;
			if (box_entry(P, ray, ex, ey, ez)) {
				tally.box_hit = 1u;
				first_hit = -2;
				AxisState ax, ay, az;
				ax.p = fadd(ex, fmul(P.nudge, ray.dx));           // main/hmap.cpp:998
				ay.p = fadd(ey, fmul(P.nudge, ray.dy));
				az.p = fadd(ez, fmul(P.nudge, ray.dz));
				ax.s = fmul(P.step_dist, ray.dx);                  // :1037
				ay.s = fmul(P.step_dist, ray.dy);
				az.s = fmul(P.step_dist, ray.dz);
				ax.tag = ay.tag = az.tag = INT_MIN;
				ax.S = ay.S = az.S = 0.0;

				// per-step motion in fixed-point units (estimates only; never decide a result)
				const float dxf = (float)(ax.s * P.fx_scale);
				const float dyf = (float)(-ay.s * P.fx_scale);
				const float dzf = (float)(az.s * P.zq_scale);
				const float inv_adx = 1.0f / fabsf(dxf), inv_ady = 1.0f / fabsf(dyf), inv_adz = 1.0f / fabsf(dzf);

				unsigned long long steps = 0ULL, fetches = 0ULL;
				int level = P.lmin + 2 * P.lstride;
				if (level > P.ltop) level = P.ltop;

				int vx, vy;
				int cx = axis_cell(ax.p, __fma_rn(ax.p, P.fx_scale, HMRM_MAGIC), k, P.gw, vx);
				int cy = axis_cell(-ay.p, __fma_rn(ay.p, -P.fx_scale, HMRM_MAGIC), k, P.gw, vy);
				int zq = zq_of(az.p, P.zq_scale, P.zq_offset);

				for (;;) {
					if (cx < 0 || cy < 0 || cx >= P.map_w || cy >= P.map_h) break;   // :1006-1011

					bool stepped = false;
					while (level >= P.lmin) {
						const int shift = k + level;
						const int bx = cx >> level, by = cy >> level;
						const int qmax = (int)__ldg(P.mip[level] + (size_t)by * (size_t)P.mip_w[level] + (size_t)bx);
						fetches += 1ULL;
						if (zq <= qmax) {
							level -= P.lstride;      // not clear of this block: look closer
							continue;
						}
						// Sample is above everything in block (bx,by): it cannot hit.  How far can we go?
						float est = (float)HMRM_JUMP_CAP;
						if (vx != INT_MAX && vx != INT_MIN) {
							const int edge = (dxf > 0.0f) ? (((bx + 1) << shift) - vx) : (vx - (bx << shift));
							est = fminf(est, __int2float_rz(edge) * inv_adx);
						}
						if (vy != INT_MAX && vy != INT_MIN) {
							const int edge = (dyf > 0.0f) ? (((by + 1) << shift) - vy) : (vy - (by << shift));
							est = fminf(est, __int2float_rz(edge) * inv_ady);
						}
						if (dzf < 0.0f) est = fminf(est, __int2float_rz(zq - qmax) * inv_adz);
						int m = (est >= 2.0f) ? __float2int_rz(est) : 1;
						if (m >= 2) {
							// evaluate all three (each may need its increment refreshed)
							const bool rx = axis_ready(ax), ry = axis_ready(ay), rz = axis_ready(az);
							if (!(rx && ry && rz)) m = 1;
						}

						while (m >= 2) {
							const double md = small_int_to_double(m);
							double nx, ny, nz;
							if (axis_jump(ax, md, nx) && axis_jump(ay, md, ny) && axis_jump(az, md, nz)) {
								// a ray that cannot move sideways and is not coming down never ends (reference hangs)
								if (ax.S == 0.0 && ay.S == 0.0 && !(az.S < 0.0)) {
									tally.cut_off = 1u;
									steps += 1ULL;
									m = -1;
									break;
								}
								int wx, wy;
								const bool okx = magic_decode(__fma_rn(nx, P.fx_scale, HMRM_MAGIC), wx);
								const bool oky = magic_decode(__fma_rn(ny, -P.fx_scale, HMRM_MAGIC), wy);
								const int wz = zq_of(nz, P.zq_scale, P.zq_offset);
								// (3): end point inside the same block (with the +-1 unit uncertainty), in the grid, above the block
								if (okx && oky && wx >= 1 && wy >= 1 && ((wx - 1) >> shift) == bx && ((wx + 1) >> shift) == bx &&
								    ((wy - 1) >> shift) == by && ((wy + 1) >> shift) == by && ((wx + 1) >> k) < P.map_w &&
								    ((wy + 1) >> k) < P.map_h && wz > qmax) {
									ax.p = nx; ay.p = ny; az.p = nz;
									steps += (unsigned long long)m;
									int dummy;
									cx = axis_cell(nx, __fma_rn(nx, P.fx_scale, HMRM_MAGIC), k, P.gw, dummy);
									cy = axis_cell(-ny, __fma_rn(ny, -P.fx_scale, HMRM_MAGIC), k, P.gw, dummy);
									vx = wx; vy = wy; zq = wz;
									break;
								}
							}
							m >>= 1;
						}
						if (m < 0) break;   // cut off
						if (m >= 2) {
							// jumped; the end point is the next sample to examine, one level coarser
							level += P.lstride;
							if (level > P.ltop) level -= P.lstride;
						}
						else {
							// plain single step (always exact), no cell test needed for the current sample
							const double nx = fadd(ax.p, ax.s), ny = fadd(ay.p, ay.s), nz = fadd(az.p, az.s);
							if (nx == ax.p && ny == ay.p && !(nz < az.p)) {
								tally.cut_off = 1u;
								steps += 1ULL;
								m = -1;
								break;
							}
							ax.p = nx; ay.p = ny; az.p = nz;
							steps += 1ULL;
							cx = axis_cell(nx, __fma_rn(nx, P.fx_scale, HMRM_MAGIC), k, P.gw, vx);
							cy = axis_cell(-ny, __fma_rn(ny, -P.fx_scale, HMRM_MAGIC), k, P.gw, vy);
							zq = zq_of(nz, P.zq_scale, P.zq_offset);
						}
						stepped = true;
						break;
					}
					if (tally.cut_off) break;
					if (stepped) continue;

					// finest level: the reference's own test on this cell (main/hmap.cpp:1013-1016)
					const size_t cell = (size_t)cx + (size_t)cy * (size_t)P.map_w;
					const int q = (int)__ldg(P.q0 + cell);
					fetches += 1ULL;
					steps += 1ULL;
					bool hit = zq < q;
					if (zq == q) {
						hit = az.p < __ldg(P.surf + cell);
						fetches += 1ULL;
					}
					if (hit) {
						rgba = hit_colour(P, __ldg(P.color + cell));
						real_hit = true;
						first_hit = (steps - 1ULL > 0x7FFFFFFFULL) ? 0x7FFFFFFF : (int)(steps - 1ULL);
						break;
					}
					const double nx = fadd(ax.p, ax.s), ny = fadd(ay.p, ay.s), nz = fadd(az.p, az.s);
					if (nx == ax.p && ny == ay.p && !(nz < az.p)) {
						tally.cut_off = 1u;
						break;
					}
					ax.p = nx; ay.p = ny; az.p = nz;
					cx = axis_cell(nx, __fma_rn(nx, P.fx_scale, HMRM_MAGIC), k, P.gw, vx);
					cy = axis_cell(-ny, __fma_rn(ny, -P.fx_scale, HMRM_MAGIC), k, P.gw, vy);
					zq = zq_of(nz, P.zq_scale, P.zq_offset);
					level = P.lmin;
				}
				tally.steps = steps;
				tally.fetches = fetches;
			}
			if (!real_hit) rgba = miss_colour(P, ray.dz);
			else tally.surf_hit = 1u;
			P.fb[(size_t)py * (size_t)P.W + (size_t)px] = rgba;
			if (P.step_index) P.step_index[(size_t)py * (size_t)P.W + (size_t)px] = first_hit;
		}
		commit_tally<kStats>(P, active, tally);
	}
}

} // namespace hmrm

#endif
