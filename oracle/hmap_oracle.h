/* TEST INFRASTRUCTURE — not product code.
 *
 * CPU oracle: a plain-C restatement of the reference's per-pixel heightmap
 * ray-march path (Costava/heightmap-ray-marcher, main/hmap.cpp:171-191,
 * :659-672, :952-1058; src/Perspective.cpp, src/Spherical.cpp,
 * src/Orthographic.cpp, src/AABB.cpp).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Pinning: the reference ships no tests or golden vectors, so this oracle is
 * pinned against the UNMODIFIED reference compiled here (oracle/_ref/hmap_ref,
 * oracle/_ref/ref_harness; recipe in oracle/Makefile) — frames, heights, rays
 * and AABB distances are compared bit for bit in tests/test_oracle_vs_ref.py,
 * and the committed fixtures under tests/golden/ were produced by that
 * reference build (tests/golden/make_golden.py).
 */
#ifndef HMRM_ORACLE_H
#define HMRM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One frame's worth of the reference's global render state (main/hmap.cpp:28-112). */
typedef struct oracle_frame {
	int32_t projection;      /* 1 perspective, 2 spherical, 3 orthographic (:104-107) */
	int32_t screen_width;    /* :31 */
	int32_t screen_height;   /* :32 */
	double cam_pos[3];       /* :75 */
	double hang;             /* radians, :80 */
	double vang;             /* radians, :85 */
	double hfov;             /* radians, :35 */
	double ortho_width;      /* :98 */
	double grid_width;       /* :65 */
	double step_dist;        /* :68 */
	double min_height;       /* :38 */
	double max_height;       /* :39 */
	uint8_t bg[3];           /* :110-112 */
	uint8_t pad_;
	int32_t cycle;           /* first pixel index rendered (:979) */
	int32_t cycle_period;    /* pixel stride (:980) */
} oracle_frame;

typedef struct oracle_stats {
	int64_t rays;        /* pixels visited */
	int64_t box_hits;    /* rays whose AABB test passed */
	int64_t surf_hits;   /* rays that hit the terrain */
	int64_t steps;       /* height fetches = loop iterations reaching main/hmap.cpp:1013 */
	int64_t max_steps;   /* longest single ray */
	int32_t status;      /* 0 ok, 1 = iteration cap reached (reference would not terminate) */
} oracle_stats;

double oracle_deg2rad(double deg);                                   /* main/hmap.cpp:131-133 */

void oracle_update_heightmap(const uint8_t *rgb8, int64_t num_pixels, /* main/hmap.cpp:171-191 */
                             double lum_r, double lum_g, double lum_b,
                             double min_height, double max_height, double *heights);

/* main/hmap.cpp:661-672 */
void oracle_camera_basis(double hang, double vang, double look[3], double up[3]);

/* ImagePlane::GetRay for the three projections (src/ of the reference); w,h in [0,1]. */
void oracle_get_ray(const oracle_frame *f, double w, double h, double pos[3], double dir[3]);

/* src/AABB.cpp:49-77 and :30-47 */
double oracle_distance(const double pos[3], const double dir[3], const double c0[3], const double c1[3]);
int oracle_intersection(double out[3], const double pos[3], const double dir[3],
                        const double c0[3], const double c1[3]);

/* main/hmap.cpp:952-1058.  heights: double[map_h][map_w]; colormap: RGBA8;
 * framebuf: RGBA8 [H][W][4] (only the pixels of this cycle phase are written);
 * step_index (optional): per pixel, index of the first-hit sample, -1 = box
 * missed, -2 = box entered but no surface hit; row_begin/row_end restrict the
 * rendered rows (for sampled CPU baselines; pass 0, screen_height for all). */
int oracle_render(const oracle_frame *f, const double *heights, const uint8_t *colormap,
                  int32_t map_w, int32_t map_h, uint8_t *framebuf, int32_t *step_index,
                  int32_t row_begin, int32_t row_end, oracle_stats *stats);

/* The same loop over the procedural synthetic map (csrc/synth_fbm.h) of size 2^log2n, texels generated on the fly:
 * for map sizes whose FP64 height array does not fit the test host (BASELINE configs 3-5: 8192^2 .. 32768^2).
 * Heights are UpdateHeightmap's expression (main/hmap.cpp:171-191) evaluated for the texel's grey value. */
int oracle_render_synth(const oracle_frame *f, uint32_t log2n, uint32_t seed, double lum_r, double lum_g, double lum_b,
                        uint8_t *framebuf, int32_t *step_index, int32_t row_begin, int32_t row_end,
                        oracle_stats *stats);

/* synthetic maps (csrc/synth_fbm.h) on the CPU: rgb8 [n][n][3], rgba8 [n][n][4] */
void oracle_synth_maps(uint32_t log2n, uint32_t seed, uint8_t *height_rgb8, uint8_t *color_rgba8);

#ifdef __cplusplus
}
#endif

#endif
