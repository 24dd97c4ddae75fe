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

// How many whole steps can this axis take (closed form) before it leaves its binade?  Estimate only: the
// caller verifies the end point.  |p| in [2^e, 2^(e+1)).
__device__ __forceinline__ int steps_to_binade_edge(const AxisState &a) {
	const double lo_edge = __hiloint2double(__double2hiint(a.p) & 0x7FF00000, 0);   // 2^e
	const double ap = fabs(a.p), aS = fabs(a.S);
	const bool toward_zero = (a.S < 0.0) == (a.p > 0.0);
	const double room = toward_zero ? fsub(ap, lo_edge) : fsub(fmul(lo_edge, 2.0), ap);
	// an FP32 quotient is enough for an estimate (an IEEE double divide costs ~13 DADD issue slots); erring low only
	// costs another round of the caller's loop
	float inv;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(__double2float_rn(aS)));
	const float r = __double2float_rz(room) * inv * 0.999f;
	if (!(r < (float)HMRM_JUMP_CAP)) return HMRM_JUMP_CAP;
	const int n = __float2int_rz(r);
	return toward_zero ? n : n - 1;
}

// Fixed-point cell coordinate of one axis.  t = fma(c, 2^k/gw, magic) where c is x or -y.
// Returns true when the low word v is a valid non-negative decode that is at least one unit away from a
// cell edge; then |v - 2^k*fl(c/gw)| < 1 implies (int)(c/gw) == v >> k  (main/hmap.cpp:1001-1004).
__device__ __forceinline__ bool fixed_axis(double t, unsigned mask, int &v) {
	v = __double2loint(t);
	return __double2hiint(t) == 0x43380000 && v >= 0 && (((unsigned)v + 1u) & mask) > 1u;
}

// A sample in the fixed-point view.  vx >> k and vy >> k ARE the reference's cell indices (exact): when the
// fixed-point value is too close to a cell edge, negative, huge or NaN, the reference's own divide decides and
// the coordinate is replaced by the centre of that cell (or by an out-of-grid marker); `fast` then is false and
// the sample may not be the end point of a jump.
struct MarchPos {
	int vx, vy;
	int zq;         // Zq(z)
	bool fast;
};

__device__ __forceinline__ int exact_axis(double c, double gw, int dim, int k) {
	const int cell = trunc_cell(fdiv(c, gw));                 // the reference's expression, main/hmap.cpp:1001-1004
	if (cell < 0) return -1;
	if (cell >= dim) return INT_MAX;
	return (cell << k) + (1 << (k - 1));
}

__device__ __forceinline__ MarchPos locate(const RenderParams &P, double x, double y, double z) {
	MarchPos r;
	const int k = P.fx_bits;
	const unsigned mask = (1u << k) - 1u;
	const bool okx = fixed_axis(__fma_rn(x, P.fx_scale, HMRM_MAGIC), mask, r.vx);
	const bool oky = fixed_axis(__fma_rn(y, -P.fx_scale, HMRM_MAGIC), mask, r.vy);
	r.fast = okx && oky;
	if (!r.fast) {
		if (!okx) r.vx = exact_axis(x, P.gw, P.map_w, k);
		if (!oky) r.vy = exact_axis(-y, P.gw, P.map_h, k);
	}
	r.zq = zq16(z, P.zq_scale, P.zq_offset);      // clamped exactly as K1 clamps the stored values
	return r;
}

__device__ __forceinline__ float fast_rcp(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

template <bool kStats>
__device__ __forceinline__ void march_skip(const RenderParams &P, const Ray &ray, double ex, double ey, double ez,
                                           uint32_t &rgba, bool &real_hit, int &first_hit, PixelTally &tally) {
	const int k = P.fx_bits;
	AxisState ax, ay, az;
	ax.p = fadd(ex, fmul(P.nudge, ray.dx));           // main/hmap.cpp:998
	ay.p = fadd(ey, fmul(P.nudge, ray.dy));
	az.p = fadd(ez, fmul(P.nudge, ray.dz));
	ax.s = fmul(P.step_dist, ray.dx);                  // :1037
	ay.s = fmul(P.step_dist, ray.dy);
	az.s = fmul(P.step_dist, ray.dz);
	ax.tag = ay.tag = az.tag = INT_MIN;
	ax.S = ay.S = az.S = 0.0;

	// per-step motion in fixed-point units: estimates only, they never decide a result
	const float adx = fabsf((float)(ax.s * P.fx_scale)), ady = fabsf((float)(ay.s * P.fx_scale));
	const float adz = fabsf((float)(az.s * P.zq_scale));
	const float inv_adx = fast_rcp(adx), inv_ady = fast_rcp(ady), inv_adz = fast_rcp(adz);
	const bool x_up = __double2hiint(ax.s) >= 0;       // vx grows with x
	const bool y_up = __double2hiint(ay.s) < 0;        // vy grows with -y
	const bool z_down = az.s < 0.0;
	const int cell_exit = (int)fminf(P.cell_exit_scale * adz + 32.0f, 1.0e9f);
	const unsigned grid_vx = (unsigned)P.map_w << k, grid_vy = (unsigned)P.map_h << k;   // <= 2^30

	// one level of the pyramid: lv_desc[l] = (element offset, pitch); level 0 is Zq(surf) per cell, levels >= 1
	// hold the maximum over the 2^l-cell block and its eight neighbours
	auto probe = [&](int level, int vx, int vy) -> int {
		const uint2 d = P.lv_desc[level];
		const unsigned idx = d.x + pyr_index_rt(P.layout, (unsigned)(vx >> (k + level)), (unsigned)(vy >> (k + level)), d.y);
		return (int)__ldg(P.lv + HMRM_CHECKED(P, idx, P.lv_total));
	};
	// extent (clipped to the grid) of the neighbourhood a level-`level` texel covers, in fixed-point units
	auto extent = [&](int level, int v, unsigned grid, int &lo, int &hi) {
		const int sh = k + level;
		const int b = v >> sh;
		lo = max(b - 1, 0) << sh;
		hi = (int)min((unsigned)(b + 2) << sh, grid);
	};

	unsigned steps = 0u, fetches = 0u, iters = 0u;
	int level = P.lstart;
	MarchPos cur = locate(P, ax.p, ay.p, az.p);

	while ((unsigned)cur.vx < grid_vx && (unsigned)cur.vy < grid_vy) {       // :1006-1011
		// ---- A: find a level whose neighbourhood this sample clears (descend), or reach the cell itself ----
		if (kStats) iters += 1u;
		int q = probe(level, cur.vx, cur.vy);
		if (kStats) fetches += 1u;
		while (cur.zq <= q && level > 0) {
			level = (level - P.lstride >= P.lmin) ? level - P.lstride : 0;
			q = probe(level, cur.vx, cur.vy);
			if (kStats) { fetches += 1u; tally.dbg[3] += 1u; }
		}

		int m = 1;
		int jl = level;          // level whose neighbourhood bounds this advance
		if (cur.zq > q) {
			// above every surface value of the neighbourhood: the sample cannot hit
			if (level > 0) {
				float est_xy, est_z;
				for (;;) {
					int lo, hi;
					extent(level, cur.vx, grid_vx, lo, hi);
					const float ex_ = __int2float_rz(x_up ? hi - cur.vx : cur.vx - lo) * inv_adx;
					extent(level, cur.vy, grid_vy, lo, hi);
					const float ey_ = __int2float_rz(y_up ? hi - cur.vy : cur.vy - lo) * inv_ady;
					est_xy = fminf(ex_, ey_);
					est_z = z_down ? __int2float_rz(cur.zq - q) * inv_adz : 3.0e38f;
					// climb while z leaves room for (much) wider blocks and the wider neighbourhood is cleared too
					if (!(est_z >= 4.0f * est_xy) || level + P.lstride > P.ltop) break;
					const int q2 = probe(level + P.lstride, cur.vx, cur.vy);
					if (kStats) fetches += 1u;
					if (cur.zq <= q2) break;
					level += P.lstride;
					q = q2;
				}
				jl = level;
				const float est = fminf(fminf(est_xy, est_z), (float)HMRM_JUMP_CAP) * 0.999f;
				m = (est >= 2.0f) ? __float2int_rz(est) : 1;
				// a z-limited jump ends just above q: the next sample will need a finer level
				if (est_z < est_xy) level = (level - P.lstride >= P.lmin) ? level - P.lstride : 0;
			}
			else {
				if (kStats) tally.dbg[4] += 1u;
				if (cur.zq - q > cell_exit) level = P.lmin;
			}
		}
		else {
			// level 0 and not above: the reference's own test on this cell (main/hmap.cpp:1013-1016)
			const size_t cell = (size_t)(cur.vx >> k) + (size_t)(cur.vy >> k) * (size_t)P.map_w;
			bool hit = cur.zq < q;
			if (kStats) tally.dbg[5] += 1u;
			if (!hit) {
				hit = az.p < __ldg(P.surf + HMRM_CHECKED(P, cell, (size_t)P.map_w * (size_t)P.map_h));     // Zq tie: decide in FP64
				if (kStats) fetches += 1u;
			}
			if (hit) {
				rgba = hit_colour(P, __ldg(P.color + HMRM_CHECKED(P, cell, (size_t)P.map_w * (size_t)P.map_h)));
				real_hit = true;
				if (kStats) first_hit = (steps > 0x7FFFFFFFu) ? 0x7FFFFFFF : (int)steps;
				steps += 1u;
				break;
			}
		}

		// ---- C: advance m samples: closed form when m >= 2 (end point verified), the reference's plain add otherwise ----
		bool jump = false;
		double md = 1.0;
		if (m >= 2) {
			const bool rx = axis_ready(ax), ry = axis_ready(ay), rz = axis_ready(az);
			jump = rx && ry && rz && cur.fast;
			if (jump && m == HMRM_JUMP_CAP && ax.S == 0.0 && ay.S == 0.0 && !(az.S < 0.0)) {
				tally.cut_off = 1u;            // cannot move sideways, not coming down: the reference never ends
				steps += 1u;
				break;
			}
			md = small_int_to_double(m);
		}
		double nx = jump ? __fma_rn(md, ax.S, ax.p) : fadd(ax.p, ax.s);
		double ny = jump ? __fma_rn(md, ay.S, ay.p) : fadd(ay.p, ay.s);
		double nz = jump ? __fma_rn(md, az.S, az.p) : fadd(az.p, az.s);
		MarchPos nxt = locate(P, nx, ny, nz);
		if (jump) {
			// (1) same binades, (3) end point inside the cleared neighbourhood (one unit clear of its edges), above q
			int lo_x, hi_x, lo_y, hi_y;
			extent(jl, cur.vx, grid_vx, lo_x, hi_x);
			extent(jl, cur.vy, grid_vy, lo_y, hi_y);
			const bool in_binade = binade_tag(nx) == ax.tag && binade_tag(ny) == ay.tag && binade_tag(nz) == az.tag;
			bool ok = in_binade && nxt.fast && nxt.vx - lo_x >= 1 && hi_x - nxt.vx >= 2 && nxt.vy - lo_y >= 1 &&
			          hi_y - nxt.vy >= 2 && nxt.zq > q;
			if (!in_binade) {
				// some axis would leave its binade: go exactly as far as the binade allows, the next plain step crosses
				int m2 = m;
				if (binade_tag(nx) != ax.tag) m2 = min(m2, steps_to_binade_edge(ax));
				if (binade_tag(ny) != ay.tag) m2 = min(m2, steps_to_binade_edge(ay));
				if (binade_tag(nz) != az.tag) m2 = min(m2, steps_to_binade_edge(az));
				if (m2 >= 2 && m2 < m) {
					m = m2;
					md = small_int_to_double(m);
					nx = __fma_rn(md, ax.S, ax.p);
					ny = __fma_rn(md, ay.S, ay.p);
					nz = __fma_rn(md, az.S, az.p);
					nxt = locate(P, nx, ny, nz);
					ok = binade_tag(nx) == ax.tag && binade_tag(ny) == ay.tag && binade_tag(nz) == az.tag && nxt.fast &&
					     nxt.vx - lo_x >= 1 && hi_x - nxt.vx >= 2 && nxt.vy - lo_y >= 1 && hi_y - nxt.vy >= 2 && nxt.zq > q;
					level = jl;
				}
			}
			if (!ok) {
				if (kStats) tally.dbg[6] += 1u;
				jump = false;
				nx = fadd(ax.p, ax.s);
				ny = fadd(ay.p, ay.s);
				nz = fadd(az.p, az.s);
				nxt = locate(P, nx, ny, nz);
				level = jl;
			}
		}
		if (!jump) {
			m = 1;
			// a sample that does not move sideways keeps its fixed-point coordinates: test those first
			if (nxt.vx == cur.vx && nxt.vy == cur.vy && nx == ax.p && ny == ay.p && !(nz < az.p)) {
				tally.cut_off = 1u;            // the reference would loop forever
				steps += 1u;
				break;
			}
		}
		if (kStats) {
			if (jump) { tally.dbg[0] += 1u; tally.dbg[1] += (unsigned)m; }
			else if (jl > 0) tally.dbg[2] += 1u;
			if (!nxt.fast) tally.dbg[7] += 1u;
		}
		ax.p = nx; ay.p = ny; az.p = nz;
		cur = nxt;
		steps += (unsigned)m;
	}
	tally.steps = steps;
	tally.fetches = fetches;
	if (kStats) atomicMax(&P.stats->dbg[10], (unsigned long long)iters);    // most loop iterations of any ray
}

template <bool kStats>
__global__ void __launch_bounds__(256, 4) k2_render_skip(const __grid_constant__ RenderParams P) {
	const int lane = threadIdx.x & 31;
	const unsigned n_tiles = (unsigned)(P.tiles_x * P.tiles_y);

	for (;;) {
		const unsigned tile = next_tile(P.tile_counter);
		if (tile >= n_tiles) break;
		int ty_seq, tx;
		tile_row_col(P, tile, ty_seq, tx);
		// longest-processing-time-first: a few grazing tiles take ~50x the mean, so they must start early
		const int ty = P.row_order ? __ldg(P.row_order + HMRM_CHECKED(P, ty_seq, P.tiles_y)) : ty_seq;
		const int px = tx * 8 + (lane & 7);
		const int py = P.row_begin + (P.tile_y_first + ty * P.tile_y_step) * 4 + (lane >> 3);
		const bool active = pixel_selected(P, px, py);
		long long t_tile = 0;
		if (kStats) t_tile = clock64();

		PixelTally tally = {0ULL, 0ULL, 0u, 0u, 0u, {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}};
		uint32_t rgba = 0u;
		if (active) {
			bool real_hit = false;
			int first_hit = -1;
			const Ray ray = generate_ray(P, px, py);
			double ex, ey, ez, lo = 0.0;
			const bool entered = box_entry(P, ray, ex, ey, ez, lo);
			if (kStats && P.ray_dump) dump_ray(P, px, py, ray, entered, lo, ex, ey, ez);
			if (entered) {
				tally.box_hit = 1u;
				first_hit = -2;
				march_skip<kStats>(P, ray, ex, ey, ez, rgba, real_hit, first_hit, tally);
			}
			if (!real_hit) rgba = miss_colour(P, ray.dz);
			else tally.surf_hit = 1u;
			if (P.step_index) P.step_index[(size_t)py * (size_t)P.W + (size_t)px] = first_hit;
		}
		store_pixel(P, px, py, active, rgba);
		commit_tally<kStats>(P, active, tally);
		if (kStats && lane == 0) {
			const unsigned long long dt = (unsigned long long)(clock64() - t_tile);
			atomicMax(&P.stats->dbg[8], dt);       // slowest tile (SM clocks)
			atomicAdd(&P.stats->dbg[9], dt);       // sum over tiles
			if (dt > 100000ULL) atomicAdd(&P.stats->dbg[11], 1ULL);   // tiles above 100 k clocks
		}
	}
}

} // namespace hmrm

#endif
