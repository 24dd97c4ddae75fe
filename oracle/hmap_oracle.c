/* TEST INFRASTRUCTURE — not product code.  See hmap_oracle.h.
 *
 * Plain-C (C99) restatement of the reference's ray-march path, written from
 * the arithmetic contract in SURVEY.md Appendix A: every floating-point
 * operation below is one IEEE-754 binary64 operation in the order in which
 * the reference's C++ expressions evaluate, so the result is bit-identical to
 * the reference when compiled without FMA contraction (-ffp-contract=off).
 * Differences from the reference: 64-bit pixel/texel indices (the reference's
 * `int` indices overflow above ~23k x 23k maps, main/hmap.cpp:177,1018),
 * optional per-pixel first-hit step index and step statistics, an iteration
 * cap instead of a hang for rays that never leave the grid.
 */
#include "hmap_oracle.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../heightmap-ray-marcher_b200/csrc/synth_fbm.h"

typedef struct v3 { double x, y, z; } v3;

static v3 v3_make(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static v3 v3_scale(double s, v3 a) { return v3_make(s * a.x, s * a.y, s * a.z); }

/* GLM 0.9.9.8 scalar definitions (see oracle/shim/glm/glm.hpp). */
static v3 v3_cross(v3 a, v3 b) {
	return v3_make(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static v3 v3_normalize(v3 a) {
	const double d = (a.x * a.x + a.y * a.y) + a.z * a.z;
	const double inv = 1.0 / sqrt(d);
	return v3_make(a.x * inv, a.y * inv, a.z * inv);
}

/* main/hmap.cpp:131-133 */
double oracle_deg2rad(double deg) { return (deg / 180.0) * M_PI; }

/* main/hmap.cpp:171-191 */
void oracle_update_heightmap(const uint8_t *rgb8, int64_t num_pixels,
                             double lum_r, double lum_g, double lum_b,
                             double min_height, double max_height, double *heights) {
	const double span = max_height - min_height;
#pragma omp parallel for schedule(static)
	for (int64_t p = 0; p < num_pixels; ++p) {
		const double r = (double)rgb8[3 * p + 0];
		const double g = (double)rgb8[3 * p + 1];
		const double b = (double)rgb8[3 * p + 2];
		double v = (lum_r * r + lum_g * g) + lum_b * b;
		if (v < 0.0) v = 0.0;
		else if (v > 255.0) v = 255.0;
		heights[p] = (v / 255.0) * span + min_height;
	}
}

/* main/hmap.cpp:661-672 */
void oracle_camera_basis(double hang, double vang, double look[3], double up[3]) {
	const double sv = sin(vang), cv = cos(vang);
	const double sh = sin(hang), ch = cos(hang);
	look[0] = sv * ch;
	look[1] = sv * sh;
	look[2] = cv;
	const double uv = vang - (M_PI / 2.0);
	const double su = sin(uv), cu = cos(uv);
	up[0] = su * ch;
	up[1] = su * sh;
	up[2] = cu;
}

/* Per-frame image-plane constants (the three constructors). */
typedef struct plane {
	int projection;
	v3 cam;
	v3 ul, pr, pd;      /* perspective / orthographic plane */
	v3 look_f;          /* orthographic: float-rounded look */
	double hfov, vfov, ul_hang, ul_vang; /* spherical */
} plane;

static plane plane_build(const oracle_frame *f) {
	plane pl;
	double lk[3], u[3];
	const double ar = (double)f->screen_width / f->screen_height;   /* main/hmap.cpp:956,960 */
	pl.projection = f->projection;
	pl.cam = v3_make(f->cam_pos[0], f->cam_pos[1], f->cam_pos[2]);
	pl.ul = pl.pr = pl.pd = pl.look_f = v3_make(0.0, 0.0, 0.0);
	pl.hfov = pl.vfov = pl.ul_hang = pl.ul_vang = 0.0;
	oracle_camera_basis(f->hang, f->vang, lk, u);

	if (f->projection == 1) {
		/* src/Perspective.cpp:3-23 */
		const v3 look = v3_make(lk[0], lk[1], lk[2]);
		const v3 up = v3_make(u[0], u[1], u[2]);
		const double hpw = tan(f->hfov / 2.0);
		const double hph = hpw / ar;
		const v3 right = v3_normalize(v3_cross(look, up));
		const v3 centre = v3_add(pl.cam, look);
		const v3 vup = v3_scale(hph, up);
		const v3 vright = v3_scale(hpw, right);
		const v3 ll = v3_sub(v3_sub(centre, vup), vright);
		const v3 ur = v3_add(v3_add(centre, vup), vright);
		pl.ul = v3_sub(v3_add(centre, vup), vright);
		pl.pr = v3_sub(ur, pl.ul);
		pl.pd = v3_sub(ll, pl.ul);
	}
	else if (f->projection == 2) {
		/* src/Spherical.cpp:3-15 */
		pl.hfov = f->hfov;
		pl.vfov = f->hfov / ar;
		pl.ul_hang = f->hang + (f->hfov / 2.0);
		pl.ul_vang = f->vang - (pl.vfov / 2.0);
	}
	else {
		/* src/Orthographic.cpp:3-17; look/up arrive as float vectors (main/hmap.cpp:963) */
		const v3 look = v3_make((double)(float)lk[0], (double)(float)lk[1], (double)(float)lk[2]);
		const v3 up = v3_make((double)(float)u[0], (double)(float)u[1], (double)(float)u[2]);
		const v3 right = v3_cross(look, up);
		const double half_w = (f->screen_width / 2.0) * f->ortho_width;
		const double half_h = (f->screen_height / 2.0) * f->ortho_width;
		pl.look_f = look;
		pl.ul = v3_add(v3_sub(pl.cam, v3_scale(half_w, right)), v3_scale(half_h, up));
		pl.pr = v3_scale(f->screen_width * f->ortho_width, right);
		pl.pd = v3_scale(f->screen_height * f->ortho_width, v3_neg(up));
	}
	return pl;
}

static void plane_ray(const plane *pl, double w, double h, v3 *pos, v3 *dir) {
	if (pl->projection == 1) {
		/* src/Perspective.cpp:25-32 */
		const v3 on_plane = v3_add(v3_add(pl->ul, v3_scale(w, pl->pr)), v3_scale(h, pl->pd));
		*pos = pl->cam;
		*dir = v3_normalize(v3_sub(on_plane, pl->cam));
	}
	else if (pl->projection == 2) {
		/* src/Spherical.cpp:17-31 */
		const double ha = pl->ul_hang - w * pl->hfov;
		const double va = pl->ul_vang + h * pl->vfov;
		const double sv = sin(va);
		*pos = pl->cam;
		*dir = v3_make(sv * cos(ha), sv * sin(ha), cos(va));
	}
	else {
		/* src/Orthographic.cpp:19-25 */
		*pos = v3_add(v3_add(pl->ul, v3_scale(w, pl->pr)), v3_scale(h, pl->pd));
		*dir = pl->look_f;
	}
}

void oracle_get_ray(const oracle_frame *f, double w, double h, double pos[3], double dir[3]) {
	const plane pl = plane_build(f);
	v3 p, d;
	plane_ray(&pl, w, h, &p, &d);
	pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
	dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

/* src/AABB.cpp:49-77 — comparison-based slab test (no fmin/fmax: NaN compares false). */
double oracle_distance(const double pos[3], const double dir[3], const double c0[3], const double c1[3]) {
	double lo = -INFINITY, hi = INFINITY;
	for (int i = 0; i < 3; ++i) {
		double near_t = (c0[i] - pos[i]) / dir[i];
		double far_t = (c1[i] - pos[i]) / dir[i];
		if (near_t > far_t) { const double t = near_t; near_t = far_t; far_t = t; }
		if (far_t < lo || near_t > hi) return INFINITY;
		if (near_t > lo) lo = near_t;
		if (far_t < hi) hi = far_t;
	}
	return (lo > hi) ? INFINITY : lo;
}

/* src/AABB.cpp:30-47 */
int oracle_intersection(double out[3], const double pos[3], const double dir[3],
                        const double c0[3], const double c1[3]) {
	const double d = oracle_distance(pos, dir, c0, c1);
	if (d == INFINITY) return 0;
	if (d < 0.0) return 0;
	for (int i = 0; i < 3; ++i) out[i] = pos[i] + d * dir[i];
	return 1;
}

/* (int)double as x86-64 cvttsd2si does it: out of range / NaN -> INT_MIN. */
static int32_t trunc_i32(double q) {
	if (!(q > -2147483649.0 && q < 2147483648.0)) return INT32_MIN;
	return (int32_t)q;
}

static uint8_t sky_channel(double v) {
	if (v < 0.0) v = 0.0;
	else if (v > 255.0) v = 255.0;
	return (uint8_t)floor(v);
}

#define ORACLE_STEP_CAP ((int64_t)1 << 31)

/* Where the march reads the map from: the arrays the reference holds (heightmap_buf, colormap_buf), or — for maps
 * too large to materialise on the test host (32768^2: 8 GiB of FP64 heights) — the procedural synthetic map itself:
 * the texel's grey value v is generated on the fly and its height comes from a 256-entry table of UpdateHeightmap's
 * expression for (v, v, v). */
typedef struct map_source {
	const double *heights;
	const uint8_t *colormap;
	int synth;
	uint32_t log2n, seed;
	double lut[256];
} map_source;

static inline double source_height(const map_source *m, int32_t gx, int32_t gy, size_t cell) {
	if (!m->synth) return m->heights[cell];
	return m->lut[hmrm_synth_height((uint32_t)gx, (uint32_t)gy, m->log2n, m->seed)];
}

static inline void source_texel(const map_source *m, int32_t gx, int32_t gy, size_t cell, uint8_t t[4]) {
	if (!m->synth) {
		const uint8_t *texel = m->colormap + cell * 4u;
		t[0] = texel[0]; t[1] = texel[1]; t[2] = texel[2]; t[3] = texel[3];
		return;
	}
	const uint32_t v = hmrm_synth_height((uint32_t)gx, (uint32_t)gy, m->log2n, m->seed);
	const uint32_t c = hmrm_synth_color(v, (uint32_t)gx, (uint32_t)gy, 1u << m->log2n);
	t[0] = (uint8_t)(c & 255u); t[1] = (uint8_t)((c >> 8) & 255u); t[2] = (uint8_t)((c >> 16) & 255u); t[3] = (uint8_t)(c >> 24);
}

/* main/hmap.cpp:952-1058 */
static int render_core(const oracle_frame *f, const map_source *src,
                       int32_t map_w, int32_t map_h, uint8_t *framebuf, int32_t *step_index,
                       int32_t row_begin, int32_t row_end, oracle_stats *stats) {
	const plane pl = plane_build(f);
	const int32_t W = f->screen_width, H = f->screen_height;
	const double gw = f->grid_width;
	/* main/hmap.cpp:967-974 */
	const double c0[3] = {0.0, 0.0, f->min_height};
	const double c1[3] = {c0[0] + map_w * gw, c0[1] - map_h * gw, f->max_height};
	const double nudge = gw * 0.01;                       /* :998 */
	const int64_t period = f->cycle_period;
	const int64_t first = (int64_t)row_begin * W;
	const int64_t last = (int64_t)row_end * W;
	/* first pixel index >= first that is congruent to cycle modulo period */
	int64_t start = f->cycle;
	if (start < first) start += ((first - start + period - 1) / period) * period;
	const int64_t count = (last > start) ? (last - start + period - 1) / period : 0;

	int64_t n_rays = 0, n_box = 0, n_surf = 0, n_steps = 0, n_max = 0;
	int capped = 0;

#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_rays, n_box, n_surf, n_steps) reduction(max : n_max) reduction(| : capped)
	for (int64_t i = 0; i < count; ++i) {
		const int64_t p = start + i * period;
		const int32_t px = (int32_t)(p % W), py = (int32_t)(p / W);
		uint8_t *out = framebuf + (size_t)p * 4u;            /* SetPixel :139-154 */
		v3 pos, dir;
		plane_ray(&pl, (double)px / (W - 1), (double)py / (H - 1), &pos, &dir);   /* :985-988 */

		const double o[3] = {pos.x, pos.y, pos.z}, d[3] = {dir.x, dir.y, dir.z};
		double e[3];
		int real_hit = 0;
		int64_t steps = 0;
		int32_t first_hit = -1;
		n_rays += 1;

		if (oracle_intersection(e, o, d, c0, c1)) {
			const double sx = f->step_dist * dir.x, sy = f->step_dist * dir.y, sz = f->step_dist * dir.z;
			double x = e[0] + nudge * dir.x, y = e[1] + nudge * dir.y, z = e[2] + nudge * dir.z;
			n_box += 1;
			first_hit = -2;
			for (;;) {
				const int32_t gx = trunc_i32((x - c0[0]) / gw);           /* :1001-1002 */
				const int32_t gy = trunc_i32(-(y - c0[1]) / gw);          /* :1003-1004 */
				if (gx < 0 || gy < 0 || gx >= map_w || gy >= map_h) break;
				const size_t cell = (size_t)gx + (size_t)gy * (size_t)map_w;
				if (z < source_height(src, gx, gy, cell) + c0[2]) {         /* :1016 */
					uint8_t texel[4];
					source_texel(src, gx, gy, cell, texel);
					if (texel[3] == 0) { out[0] = f->bg[0]; out[1] = f->bg[1]; out[2] = f->bg[2]; }
					else { out[0] = texel[0]; out[1] = texel[1]; out[2] = texel[2]; }
					out[3] = 255;
					real_hit = 1;
					first_hit = (steps > INT32_MAX) ? INT32_MAX : (int32_t)steps;
					steps += 1;
					break;
				}
				steps += 1;
				if (steps >= ORACLE_STEP_CAP) { capped = 1; break; }
				x += sx; y += sy; z += sz;                                   /* :1037 */
			}
		}

		if (!real_hit) {                                                   /* :1041-1057 */
			if (dir.z > 0.0) {
				const double zz = dir.z * dir.z;   /* std::pow(z,2) == z*z under -std=c++98 */
				out[0] = sky_channel(220.0 * zz + f->bg[0]);
				out[1] = sky_channel(240.0 * zz + f->bg[1]);
				out[2] = sky_channel(255.0 * dir.z + f->bg[2]);
			}
			else { out[0] = f->bg[0]; out[1] = f->bg[1]; out[2] = f->bg[2]; }
			out[3] = 255;
		}
		else n_surf += 1;

		if (step_index) step_index[p] = first_hit;
		n_steps += steps;
		if (steps > n_max) n_max = steps;
	}

	if (stats) {
		stats->rays = n_rays;
		stats->box_hits = n_box;
		stats->surf_hits = n_surf;
		stats->steps = n_steps;
		stats->max_steps = n_max;
		stats->status = capped;
	}
	return capped;
}

int oracle_render(const oracle_frame *f, const double *heights, const uint8_t *colormap,
                  int32_t map_w, int32_t map_h, uint8_t *framebuf, int32_t *step_index,
                  int32_t row_begin, int32_t row_end, oracle_stats *stats) {
	map_source src;
	memset(&src, 0, sizeof src);
	src.heights = heights;
	src.colormap = colormap;
	return render_core(f, &src, map_w, map_h, framebuf, step_index, row_begin, row_end, stats);
}

int oracle_render_synth(const oracle_frame *f, uint32_t log2n, uint32_t seed, double lum_r, double lum_g, double lum_b,
                        uint8_t *framebuf, int32_t *step_index, int32_t row_begin, int32_t row_end,
                        oracle_stats *stats) {
	map_source src;
	memset(&src, 0, sizeof src);
	src.synth = 1;
	src.log2n = log2n;
	src.seed = seed;
	for (int v = 0; v < 256; ++v) {
		const uint8_t rgb[3] = {(uint8_t)v, (uint8_t)v, (uint8_t)v};
		oracle_update_heightmap(rgb, 1, lum_r, lum_g, lum_b, f->min_height, f->max_height, &src.lut[v]);
	}
	const int32_t n = (int32_t)(1u << log2n);
	return render_core(f, &src, n, n, framebuf, step_index, row_begin, row_end, stats);
}

void oracle_synth_maps(uint32_t log2n, uint32_t seed, uint8_t *height_rgb8, uint8_t *color_rgba8) {
	const uint32_t n = 1u << log2n;
#pragma omp parallel for schedule(static)
	for (int64_t y = 0; y < (int64_t)n; ++y) {
		for (uint32_t x = 0; x < n; ++x) {
			const size_t p = (size_t)y * n + x;
			const uint32_t v = hmrm_synth_height(x, (uint32_t)y, log2n, seed);
			if (height_rgb8) {
				height_rgb8[3 * p + 0] = (uint8_t)v;
				height_rgb8[3 * p + 1] = (uint8_t)v;
				height_rgb8[3 * p + 2] = (uint8_t)v;
			}
			if (color_rgba8) {
				const uint32_t c = hmrm_synth_color(v, x, (uint32_t)y, n);
				color_rgba8[4 * p + 0] = (uint8_t)(c & 255u);
				color_rgba8[4 * p + 1] = (uint8_t)((c >> 8) & 255u);
				color_rgba8[4 * p + 2] = (uint8_t)((c >> 16) & 255u);
				color_rgba8[4 * p + 3] = (uint8_t)(c >> 24);
			}
		}
	}
}
