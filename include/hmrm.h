/* hmrm.h — C ABI of the B200-native heightmap ray-march path (libhmrm.so).
 *
 * The reference (Costava/heightmap-ray-marcher) has no library or FFI seam: its
 * hot path is inline in main() and parameterised by global variables
 * (main/hmap.cpp:28-112).  This header is the seam a maintainer would bind
 * instead: each entry point names the reference lines it replaces.  All
 * arguments are plain pointers, sizes and PODs; no C++ or torch types cross it.
 *
 * Conventions: every function returns 0 on success or an HMRM_ERR_* code;
 * hmrm_last_error() gives the message.  The caller owns every host buffer it
 * passes; the context owns all device memory.  A context is bound to one CUDA
 * device and must be used from one host thread at a time.  Nothing in this
 * library falls back to the CPU: without a CUDA device hmrm_create() fails.
 */
#ifndef HMRM_H
#define HMRM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMRM_ABI_VERSION 2

/* image_plane values, main/hmap.cpp:104-107 */
#define HMRM_PERSPECTIVE  1
#define HMRM_SPHERICAL    2
#define HMRM_ORTHOGRAPHIC 3

/* arithmetic mode of the render kernel */
#define HMRM_FP64_EXACT 0   /* bit-exact to the reference's RGBA8 framebuffer */
#define HMRM_FP32_FAST  1   /* accepted, and an ALIAS of HMRM_FP64_EXACT: the production traversal already marches on a
                             * 32/64-bit integer model and keeps FP64 for ray set-up and for the few samples the model
                             * cannot decide, so there is nothing left for a lossy mode to win.
                             * The north star's tolerance (>= 99.5 % identical pixels, |step delta| <= 1) is met
                             * with 100 % / 0 and is written out in the tests. */

/* traversal strategy (results are identical; this selects the kernel) */
#define HMRM_TRAVERSAL_AUTO  0
#define HMRM_TRAVERSAL_BRUTE 1   /* one height fetch per reference step (main/hmap.cpp:1000-1038 as written) */
#define HMRM_TRAVERSAL_SKIP  2   /* conservative max-mip empty-space skip in whole steps (integer linear model; default) */
#define HMRM_TRAVERSAL_SKIP_FP64 3   /* the same skip with exact FP64 positions at every decision (k2_render_skip.cuh) */
#define HMRM_TRAVERSAL_PACK  4   /* HMRM_TRAVERSAL_SKIP's traversal with lanes refilled from a per-warp stack of rays
                                  * (k2_render_pack.cuh) */

/* hmrm_frame.flags */
#define HMRM_FLAG_STATS      1u  /* count rays / steps / fetches (hmrm_get_stats) */
#define HMRM_FLAG_STEP_INDEX 2u  /* record the per-pixel first-hit step index (hmrm_get_step_index) */
#define HMRM_FLAG_RAY_DUMP   4u  /* record the device's ray / slab distance / entry point per pixel (hmrm_get_ray_dump) */

/* hmrm_frame.pixel_format: layout of the frame the render calls write */
#define HMRM_PIXEL_RGBA8 0       /* the reference's framebuf: [H][W][4], A = 255 (main/hmap.cpp:139-154) */
#define HMRM_PIXEL_RGB8  1       /* the same pixels without the constant alpha byte: [H][W][3] (25 % fewer bytes over
                                  * PCIe / NVLink; the north star's "RGB8 bands") */

/* memory layout of the 16-bit height pyramid the march fetches from (csrc/pyramid_layout.cuh); results are identical */
#define HMRM_LAYOUT_ROWMAJOR 0
#define HMRM_LAYOUT_TILE4    1   /* 4x4-texel (32-byte sector) tiles, tiles in row-major order */
#define HMRM_LAYOUT_ZORDER   2   /* 4x4-texel tiles in Z-order (Morton) inside 64x64-texel blocks */
#define HMRM_LAYOUT_DEFAULT  HMRM_LAYOUT_ROWMAJOR   /* measured: profiles/r02_layout_ab.txt — no layout wins */

#define HMRM_OK                  0
#define HMRM_ERR_INVALID         1   /* bad argument */
#define HMRM_ERR_CUDA            2   /* CUDA runtime error (message has the detail) */
#define HMRM_ERR_STATE           3   /* call order: maps / heightmap not set */
#define HMRM_ERR_NONTERMINATING  4   /* a ray can never leave the grid: the reference would hang (main/hmap.cpp:1000) */

typedef struct hmrm_ctx hmrm_ctx;

/* The reference's per-frame render state (globals, main/hmap.cpp:28-112), as one POD.
 * Angles are radians, as the reference stores them after DegreesToRads (:131-133). */
typedef struct hmrm_frame {
	int32_t projection;     /* image_plane :107 */
	int32_t screen_width;   /* :31 */
	int32_t screen_height;  /* :32 */
	int32_t precision;      /* HMRM_FP64_EXACT | HMRM_FP32_FAST */
	double cam_pos[3];      /* :75 */
	double hang;            /* :80 */
	double vang;            /* :85 */
	double hfov;            /* :35 */
	double ortho_width;     /* :98 */
	double grid_width;      /* :65 */
	double step_dist;       /* :68 */
	uint8_t bg[3];          /* bg_r, bg_g, bg_b :110-112 */
	uint8_t pixel_format;   /* HMRM_PIXEL_RGBA8 (0, the reference's framebuf) | HMRM_PIXEL_RGB8 */
	int32_t cycle;          /* first pixel index of this frame (:979), 0 <= cycle < cycle_period */
	int32_t cycle_period;   /* pixel stride (:71,:980); 1 = full frame */
	int32_t row_begin;      /* rows [row_begin,row_end) are rendered: row band of a multi-GPU frame */
	int32_t row_end;        /* 0,0 = all rows */
	int32_t traversal;      /* HMRM_TRAVERSAL_* */
	uint32_t flags;         /* HMRM_FLAG_* */
	int32_t band_count;     /* > 1: the 4-row tile rows of [row_begin,row_end) are dealt round-robin to band_count */
	int32_t band_index;     /*      renderers and this call renders those with (tile_row % band_count) == band_index */
} hmrm_frame;

typedef struct hmrm_stats {
	int64_t rays;           /* pixels rendered */
	int64_t box_hits;       /* rays that entered the AABB (src/AABB.cpp:30) */
	int64_t surf_hits;      /* rays that hit the terrain (main/hmap.cpp:1016) */
	int64_t steps;          /* reference-equivalent march steps = height fetches the reference performs (:1013) */
	int64_t fetches;        /* memory fetches this kernel actually issued on the march (any level) */
	int64_t max_steps;      /* longest single ray, in reference steps */
	int32_t status;         /* 0, or HMRM_ERR_NONTERMINATING if some ray was cut off */
	int32_t reserved0;
	double kernel_ms;       /* CUDA-event duration of the render kernel */
} hmrm_stats;

/* ---- lifetime -------------------------------------------------------------------------- */
int hmrm_abi_version(void);
int hmrm_device_count(void);
int hmrm_create(int device, hmrm_ctx **out);
void hmrm_destroy(hmrm_ctx *ctx);
const char *hmrm_last_error(const hmrm_ctx *ctx);   /* ctx may be NULL: last creation error */

/* ---- maps: replaces stbi_load results at main/hmap.cpp:320-321 (RGB8) and :341-342 (RGBA8) ---- */
int hmrm_set_maps(hmrm_ctx *ctx, const uint8_t *height_rgb8, const uint8_t *color_rgba8,
                  int32_t map_width, int32_t map_height);
/* same, from device memory of ctx's device (copied) */
int hmrm_set_maps_device(hmrm_ctx *ctx, const void *d_height_rgb8, const void *d_color_rgba8,
                         int32_t map_width, int32_t map_height);
/* synthetic fBm maps of size 2^log2n generated on the device (csrc/synth_fbm.h; bench / test input) */
int hmrm_synth_maps(hmrm_ctx *ctx, uint32_t log2n, uint32_t seed);
int hmrm_get_maps(hmrm_ctx *ctx, uint8_t *height_rgb8, uint8_t *color_rgba8);   /* download (either may be NULL) */

/* layout of the height pyramid (HMRM_LAYOUT_*; also the HMRM_LAYOUT environment variable at hmrm_create:
 * rowmajor | tile4 | zorder).  Takes effect at the next hmrm_update_heightmap, which must follow. */
int hmrm_set_layout(hmrm_ctx *ctx, int layout);
int hmrm_get_layout(hmrm_ctx *ctx);

/* ---- prepass: replaces UpdateHeightmap, main/hmap.cpp:171-191 (kernel K1) ---- */
int hmrm_update_heightmap(hmrm_ctx *ctx, const double lum[3], double min_height, double max_height);
/* heights exactly as the reference's heightmap_buf (double[H][W], main/hmap.cpp:53), for parity checks */
int hmrm_get_heights(hmrm_ctx *ctx, double *heights);

/* ---- render: replaces main/hmap.cpp:952-1058 (ImagePlane ctor + pixel loop; kernel K2) ---- */
void hmrm_frame_defaults(hmrm_frame *f);            /* the reference's defaults, :31-112, cycle_period 1 */
/* rgba_out: host RGBA8 [screen_height][screen_width][4], stride W*4 — the reference's framebuf (:612) — or, with
 * pixel_format = HMRM_PIXEL_RGB8, [screen_height][screen_width][3], stride W*3.
 * Only the pixels selected by cycle/cycle_period and the row band are written. Synchronous. */
int hmrm_render(hmrm_ctx *ctx, const hmrm_frame *f, uint8_t *rgba_out);
/* device output (same layout) on `stream` (a cudaStream_t, NULL = ctx's own stream); asynchronous */
int hmrm_render_device(hmrm_ctx *ctx, const hmrm_frame *f, void *d_rgba_out, void *stream);
/* render into the context's device framebuffer(s) and copy rows [row_begin,row_end) to (pinned or registered) host
 * memory asynchronously; hmrm_wait() joins.  rgba_out always addresses the WHOLE frame; with band_count > 1 only the
 * tile rows of this band are written (so several contexts / ranks can fill one host frame). */
int hmrm_render_async(hmrm_ctx *ctx, const hmrm_frame *f, uint8_t *rgba_out);
int hmrm_wait(hmrm_ctx *ctx);
/* Streaming: up to four whole frames (cycle_period 1) may be in flight; the copy-out of one overlaps the kernels
 * of the next ones.  hmrm_wait_pending(ctx, n) (n = 1..3) returns when every frame but the n most recent is complete
 * in its host buffer (so n + 1 host buffers are rotated); hmrm_wait_pending(ctx, 0) == hmrm_wait. */
int hmrm_wait_pending(hmrm_ctx *ctx, int max_pending);

int hmrm_get_stats(hmrm_ctx *ctx, hmrm_stats *out);            /* of the last render */
/* traversal diagnostics of the last HMRM_FLAG_STATS render with a skip traversal: [0] jumps, [1] samples covered
 * by jumps, [2] single steps above a mip block, [3] level descents, [4] cell tests above the cell, [5] cell tests
 * below the quantised height (hits), [6] HMRM_TRAVERSAL_SKIP: warp-level loop iterations (lane utilisation of the
 * march loop = ([0]+[2]+[4]+[5]+[7]) / 32 / [6]); HMRM_TRAVERSAL_SKIP_FP64: refused jumps, [7] samples decided by the
 * exact FP64 expressions, [8]..[11] HMRM_TRAVERSAL_SKIP: warp-level iterations with 1-8 / 9-16 / 17-24 / 25-32 busy
 * lanes; HMRM_TRAVERSAL_SKIP_FP64: slowest 8x4 tile in SM clocks, sum of tile clocks, most loop iterations of any ray,
 * tiles above 100 k clocks */
int hmrm_get_debug_counters(hmrm_ctx *ctx, int64_t out[12]);
int hmrm_get_step_index(hmrm_ctx *ctx, int32_t *step_index);   /* int32 [H][W]: first-hit sample index,
                                                                  -1 box missed, -2 no surface hit */
/* of the last HMRM_FLAG_RAY_DUMP render: double [H][W][10] = what the DEVICE computed for each pixel: ray pos[3] and
 * dir[3] (ImagePlane::GetRay, src/Perspective.cpp:25-32 etc.), the slab entry distance (distance(), src/AABB.cpp:49-77;
 * -inf..inf as the reference leaves it when a slab test fails early), and the entry point (intersection(), :30-47; zeros
 * when the box is missed).  Pixels not rendered by that frame are NaN. */
int hmrm_get_ray_dump(hmrm_ctx *ctx, double *out);

/* ---- type-level API mirror (host side; no device work) ---- */
/* DegreesToRads, main/hmap.cpp:131-133 */
double hmrm_deg2rad(double degrees);
/* look/up from hang,vang, main/hmap.cpp:661-672 */
void hmrm_camera_basis(double hang, double vang, double look[3], double up[3]);
/* ImagePlane::GetRay (src/ImagePlane.hpp:10) of the plane the frame describes; w,h in [0,1] */
int hmrm_get_ray(const hmrm_frame *f, double w, double h, double pos[3], double dir[3]);

/* The device's own slab test (the function the render kernels call; src/AABB.cpp:30-77) on n caller-supplied rays
 * (pos[3], dir[3] each) and boxes (c0[3], c1[3] each): out[n][5] = distance(), intersection() as 0/1, entry point.
 * For known-answer tests against the reference's functions. */
int hmrm_debug_aabb(hmrm_ctx *ctx, int32_t n, const double *rays, const double *boxes, double *out);

/* pinned host memory helpers (so that callers without a CUDA binding can stage buffers) */
int hmrm_host_alloc(void **ptr, size_t bytes);
/* the same with HMRM_HOST_* flags; write-combined memory is fast to fill from the device and slow to read on the CPU */
#define HMRM_HOST_WRITE_COMBINED 1u
int hmrm_host_alloc_flags(void **ptr, size_t bytes, uint32_t flags);
void hmrm_host_free(void *ptr);
/* page-lock memory the caller owns (e.g. a POSIX shared-memory frame that several ranks fill with their bands) */
int hmrm_host_register(void *ptr, size_t bytes);
int hmrm_host_unregister(void *ptr);

/* ---- peer frames: one big frame rendered by several GPUs of one node (one process per GPU) ----
 * The root rank allocates the frame on its device and exports a 64-byte handle (CUDA IPC); every other rank opens it
 * and passes the mapped pointer to hmrm_render_device together with band_count / band_index: its kernel then stores
 * its tile rows straight into the root's frame over NVLink — no pack, no gather, no unpack; the only collective left
 * is the barrier that tells the root the frame is complete.  New surface: the reference has one OpenMP loop
 * (main/hmap.cpp:978) and no multi-GPU path. */
#define HMRM_IPC_HANDLE_BYTES 64
int hmrm_device_alloc(hmrm_ctx *ctx, size_t bytes, void **dptr);          /* zero-filled */
int hmrm_device_free(hmrm_ctx *ctx, void *dptr);
int hmrm_ipc_export(hmrm_ctx *ctx, void *dptr, uint8_t handle[HMRM_IPC_HANDLE_BYTES]);   /* dptr from hmrm_device_alloc */
int hmrm_ipc_open(hmrm_ctx *ctx, const uint8_t handle[HMRM_IPC_HANDLE_BYTES], void **dptr); /* in ANOTHER process */
int hmrm_ipc_close(hmrm_ctx *ctx, void *dptr);

/* Device-side completion of a peer frame (csrc/peer_sync.cuh): no collective, no host synchronisation.
 * The root allocates frame bytes + HMRM_PEER_CTRL_BYTES with hmrm_device_alloc (zero-filled); the control block is the
 * last HMRM_PEER_CTRL_BYTES of it (any 128-byte aligned place will do).  `use` counts the uses of ONE buffer, from 1.
 *   every rank:  hmrm_render_peer(ctx, f, d_frame, d_ctrl, use, stream)   waits (on the device) until the root has
 *                released use - 1, renders this rank's bands (f->band_count / band_index) into d_frame and counts the
 *                rank as arrived;
 *   the root:    hmrm_peer_wait(ctx, d_ctrl, use, ranks, stream)          stream-ordered: what follows on `stream`
 *                sees the complete frame;  hmrm_peer_release(ctx, d_ctrl, use, stream) when it has been read.
 * Two implementations of the same protocol on the same block, chosen once per context (hmrm_create):
 *   stream memory operations (the default when the driver has cuStreamWaitValue32 / cuStreamWriteValue32): the GPU's
 *     front end waits and writes, no SM is involved — a one-thread kernel needs a CTA slot, and the persistent render
 *     kernels of the frames in flight hold them all (measured, 8 GPUs: 0.2006 vs 0.2181 ms per 8K frame).  Every rank
 *     writes its own arrival word.  No timeout: a missing peer leaves the stream blocked (cudaStreamQuery says so);
 *   one-thread kernels (HMRM_PEER_SYNC=kernels in the environment): `arrived` is one counter; a wait that lasts
 *     longer than 5 s (HMRM_PEER_TIMEOUT_MS) sets the block's error word instead of hanging the GPU.
 * Every rank of a frame must use the same one.  hmrm_peer_status returns {arrived counter, released, error}. */
#define HMRM_PEER_CTRL_BYTES 256
int hmrm_render_peer(hmrm_ctx *ctx, const hmrm_frame *f, void *d_frame, void *d_ctrl, uint32_t use, void *stream);
/* The same with the exchange done by the copy engine: the rank renders its bands into d_stage (a frame-sized buffer
 * on ITS device), then pushes exactly its tile rows into d_frame with one strided device-to-device copy, then counts
 * itself in.  One more pass over the rank's rows, but NVLink sees large transfers instead of the kernel's 24- / 32-byte
 * row pieces: the faster exchange when many ranks write into one root. */
int hmrm_render_peer_staged(hmrm_ctx *ctx, const hmrm_frame *f, void *d_stage, void *d_frame, void *d_ctrl, uint32_t use,
                            void *stream);
int hmrm_peer_wait(hmrm_ctx *ctx, void *d_ctrl, uint32_t use, int32_t ranks, void *stream);
int hmrm_peer_release(hmrm_ctx *ctx, void *d_ctrl, uint32_t use, void *stream);
int hmrm_peer_status(hmrm_ctx *ctx, void *d_ctrl, uint32_t out[3]);
#define HMRM_PEER_SYNC_KERNELS 0
#define HMRM_PEER_SYNC_MEMOPS 1
int hmrm_peer_sync_mode(const hmrm_ctx *ctx);     /* which implementation this context uses; < 0: ctx is NULL */

#ifdef __cplusplus
}
#endif

#endif
