#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — generates the reference's main/hmap.cpp with the three call sites of INTEGRATION.md replaced by
calls into libhmrm.so, so that the drop-in claim is tested on the reference's OWN program: same SDL loop, same console
and config parser, same recording code; only the height prepass and the per-pixel loop go to the GPU.

    python oracle/make_patched_reference.py /root/reference/main/hmap.cpp oracle/_ref/hmap_patched.cpp

The edits are addressed by line number and guarded by short anchors; no reference source text is stored in this
repository, and the generated file lands in the git-ignored oracle/_ref/.
"""
import sys

src_path, out_path = sys.argv[1], sys.argv[2]
lines = open(src_path).read().split("\n")


def expect(no: int, token: str) -> None:
    if token not in lines[no - 1]:
        raise SystemExit(f"{src_path}:{no}: expected {token!r}, found {lines[no - 1]!r}: the reference changed")


expect(171, "static void UpdateHeightmap()")
expect(191, "}")
expect(321, "&heightmap_width")
expect(342, "&colormap_width")
expect(612, "framebuf = new Uint8[")
expect(698, "delete[] framebuf;")
expect(699, "framebuf = new Uint8[")
expect(1159, "delete[] framebuf;")
expect(667, "up_vang")
expect(672, ");")
expect(952, "ImagePlane *ip;")
expect(976, "cycle = (cycle + 1) % cycle_period;")
expect(978, "#pragma omp parallel for")
expect(1058, "}")
expect(1060, "text_surface_rerender_timer_ms")
expect(1129, "delete ip;")
expect(1161, "return 0;")

GLOBALS = r'''
// ---- INTEGRATION.md §1: the B200 path behind the C ABI ----
#include "hmrm.h"
static hmrm_ctx *g_hmrm = NULL;
static bool g_maps_dirty = true, g_heights_dirty = true;
static double g_hang_top = 0.0, g_vang_top = 0.0;   // hang/vang as look/up saw them at the top of the loop (SURVEY D-6)
static void HmrmDie(const char *what) {
	std::cerr << "hmrm: " << what << ": " << hmrm_last_error(g_hmrm) << "\n";
	std::exit(1);
}
// framebuf is page-locked while it lives, so that the copy-out of a frame runs at PCIe speed (best effort)
static void HmrmPin(Uint8 *buf) { hmrm_host_register(buf, (size_t)screen_width * (size_t)screen_height * 4); }
static void HmrmUnpin(Uint8 *buf) { hmrm_host_unregister(buf); }
'''

UPDATE_BODY = r'''	// INTEGRATION.md §2: the prepass runs on the device (kernel K1) when the next frame needs it
	g_heights_dirty = true;'''

SYNC_AND_FRAME = r'''		// ---- INTEGRATION.md §2 + §3: maps / heights on the device, then the frame ----
		if (g_hmrm == NULL && hmrm_create(0, &g_hmrm) != HMRM_OK) HmrmDie("hmrm_create");
		if (g_maps_dirty) {
			if (hmrm_set_maps(g_hmrm, base_heightmap_buf, colormap_buf, heightmap_width, heightmap_height) != HMRM_OK)
				HmrmDie("hmrm_set_maps");
			g_maps_dirty = false;
			g_heights_dirty = true;
		}
		if (g_heights_dirty) {
			const double lum[3] = {lum_r, lum_g, lum_b};
			if (hmrm_update_heightmap(g_hmrm, lum, min_height, max_height) != HMRM_OK) HmrmDie("hmrm_update_heightmap");
			g_heights_dirty = false;
		}

		cycle = (cycle + 1) % cycle_period;

		{
			hmrm_frame f;
			hmrm_frame_defaults(&f);
			f.projection = image_plane;
			f.screen_width = screen_width;
			f.screen_height = screen_height;
			f.cam_pos[0] = cam_pos.x;
			f.cam_pos[1] = cam_pos.y;
			f.cam_pos[2] = cam_pos.z;
			// look/up were computed before the events were drained; Spherical takes hang/vang as they are now
			f.hang = (image_plane == IMAGEPLANE_SPHERICAL) ? hang : g_hang_top;
			f.vang = (image_plane == IMAGEPLANE_SPHERICAL) ? vang : g_vang_top;
			f.hfov = hfov;
			f.ortho_width = ortho_width;
			f.grid_width = grid_width;
			f.step_dist = step_dist;
			f.bg[0] = bg_r;
			f.bg[1] = bg_g;
			f.bg[2] = bg_b;
			f.cycle = cycle;
			f.cycle_period = cycle_period;
			f.precision = HMRM_FP64_EXACT;
			if (hmrm_render(g_hmrm, &f, framebuf) != HMRM_OK) HmrmDie("hmrm_render");
		}'''

out = []
for no, text in enumerate(lines, start=1):
    if no == 170:
        out.append(GLOBALS)
    if 172 <= no <= 190:
        if no == 172:
            out.append(UPDATE_BODY)
        continue
    if 952 <= no <= 1058:
        if no == 952:
            out.append(SYNC_AND_FRAME)
        continue
    if no == 1129:
        continue
    if no == 698:
        out.append("\t\t\t\t\tHmrmUnpin(framebuf);")
    if no == 1159:
        out.append("\tHmrmUnpin(framebuf);")
    if no == 1161:
        out.append("\thmrm_destroy(g_hmrm);")
    out.append(text)
    if no == 612:
        out.append("\tHmrmPin(framebuf);")
    if no == 699:
        out.append("\t\t\t\t\tHmrmPin(framebuf);")
    if no in (321, 342):
        out.append("\t\t\t\tg_maps_dirty = true;")
    if no == 672:
        out.append("\t\tg_hang_top = hang;\n\t\tg_vang_top = vang;")
open(out_path, "w").write("\n".join(out))
