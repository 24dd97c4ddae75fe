"""GPU: pyramid layouts, RGB8 frames, device-side intermediates against the reference's known answers, oracle parity
bands at the BASELINE map sizes, in-flight bookkeeping of the async API.  Everything goes through the C ABI."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import helpers as H
import scenes as S

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
LAYOUTS = [0, 1, 2]          # HMRM_LAYOUT_ROWMAJOR, _TILE4, _ZORDER


def load_bench():
    spec = importlib.util.spec_from_file_location("hmrm_bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


# ---------------------------------------------------------------------------------------------------------------
# N1: memory layout of the height pyramid never changes a result (same frames, step indices, steps AND fetches)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("traversal", [2, 3, 4])
@pytest.mark.parametrize("name", ["persp_graze", "spher_wide", "ortho_fine", "noise_lum_neg", "crop_ortho",
                                  "defaults_grid", "from_below", "tiny_res"])
def test_pyramid_layouts_bit_exact(hmrm, oracle, name, traversal):
    scene = S.SCENE_BY_NAME[name]
    maps = H.load_scene_maps(scene, oracle)
    ofb, osteps, ost = H.oracle_render_scene(oracle, scene, maps)
    want = H.golden_frames()[name]
    r = hmrm.Renderer(0)
    try:
        fetches = []
        for layout in LAYOUTS:
            r.set_layout(layout)
            assert r.get_layout() == layout
            H.configure(r, scene, maps)
            f = H.product_frame(hmrm, r, scene, traversal=traversal, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
            got = r.render(f)
            st = r.stats()
            assert np.array_equal(got, want), f"layout {layout}: frame differs from the reference's"
            assert np.array_equal(got, ofb) and np.array_equal(r.step_index(f), osteps)
            assert (st.steps, st.surf_hits, st.box_hits, st.status) == (ost.steps, ost.surf_hits, ost.box_hits, 0)
            fetches.append(st.fetches)
        assert len(set(fetches)) == 1, f"the traversal itself must not depend on the layout: {fetches}"
    finally:
        r.close()


def test_layout_switch_needs_a_rebuilt_pyramid(hmrm, oracle):
    scene = S.SCENE_BY_NAME["persp_basic"]
    maps = H.load_scene_maps(scene, oracle)
    r = hmrm.Renderer(0)
    try:
        H.configure(r, scene, maps)
        lib = hmrm.load_library()
        assert lib.hmrm_set_layout(r._h, 7) == 1                       # HMRM_ERR_INVALID
        other = (r.get_layout() + 1) % 3
        assert lib.hmrm_set_layout(r._h, other) == 0
        f = H.product_frame(hmrm, r, scene)
        with pytest.raises(hmrm.HmrmError) as e:                       # HMRM_ERR_STATE until hmrm_update_heightmap
            r.render(f)
        assert e.value.code == 3
        r.update_heightmap()
        assert np.array_equal(r.render(f), H.golden_frames()["persp_basic"])
    finally:
        r.close()


# ---------------------------------------------------------------------------------------------------------------
# RGB8 frames = the RGBA8 frame without its constant alpha byte (main/hmap.cpp:139-154 always writes A = 255)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("traversal", [1, 2, 3, 4])
@pytest.mark.parametrize("name", ["persp_basic", "spher_wide", "ortho_fine", "defaults_grid", "tiny_res", "alpha_zero_centre"])
def test_rgb8_frame_is_the_rgba8_frame_without_alpha(hmrm, renderer, oracle, name, traversal):
    scene = S.SCENE_BY_NAME[name]          # widths 320, 384, 256 (word stores), 333 and 9 (byte stores)
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    want = H.golden_frames()[name]
    got = renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal, pixel_format=hmrm.PIXEL_RGB8))
    assert got.shape == want.shape[:2] + (3,)
    assert np.array_equal(got, want[..., :3])


@pytest.mark.parametrize("width,height", [(320, 180), (324, 181), (322, 187)])
def test_rgb8_bands_cycle_and_device_output(hmrm, renderer, oracle, width, height):
    """Row bands, interleaved bands and the progressive interleave in RGB8; device output and async copy-out."""
    import torch

    from heightmap_ray_marcher_b200 import binding

    scene = dict(S.SCENE_BY_NAME["persp_graze"], width=width, height=height)
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    full = renderer.render(H.product_frame(hmrm, renderer, scene)).copy()[..., :3]
    rgb = dict(pixel_format=hmrm.PIXEL_RGB8)
    # contiguous row bands
    out = np.zeros_like(full)
    for a, b in [(0, 57), (57, 58), (58, height)]:
        band = np.zeros_like(full)
        renderer.render(H.product_frame(hmrm, renderer, scene, row_begin=a, row_end=b, **rgb), out=band)
        assert not band[:a].any() and not band[b:].any()
        out[a:b] = band[a:b]
    assert np.array_equal(out, full)
    # interleaved bands filling one pinned host frame
    host = binding.pinned_empty(full.shape)
    host[:] = 0
    for rk in range(3):
        renderer.render_async(H.product_frame(hmrm, renderer, scene, band_count=3, band_index=rk, **rgb), host)
        renderer.wait()
    assert np.array_equal(host, full)
    # device output
    dev = torch.zeros((height, width, 3), dtype=torch.uint8, device="cuda:0")
    renderer.render_device(H.product_frame(hmrm, renderer, scene, **rgb), dev)
    renderer.wait()
    assert np.array_equal(dev.cpu().numpy(), full)
    # progressive interleave accumulates in the persistent framebuffer
    renderer.render(H.product_frame(hmrm, renderer, scene, screen_width=64, screen_height=32, **rgb))   # fresh buffer
    for phase in range(5):
        acc = renderer.render(H.product_frame(hmrm, renderer, scene, cycle=phase, cycle_period=5, **rgb))
    assert np.array_equal(acc, full)


def test_pixel_format_switch_starts_from_a_cleared_framebuffer(hmrm, renderer, oracle):
    scene = S.SCENE_BY_NAME["persp_basic"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    renderer.render(H.product_frame(hmrm, renderer, scene))
    part = renderer.render(H.product_frame(hmrm, renderer, scene, cycle=1, cycle_period=4, pixel_format=hmrm.PIXEL_RGB8))
    sel = np.zeros(part.shape[0] * part.shape[1], dtype=bool)
    sel[1::4] = True
    sel = sel.reshape(part.shape[:2])
    assert not part[~sel].any() and np.array_equal(part[sel], H.golden_frames()["persp_basic"][..., :3][sel])
    bad = H.product_frame(hmrm, renderer, scene, pixel_format=9)
    with pytest.raises(hmrm.HmrmError) as e:
        renderer.render(bad)
    assert e.value.code == 1                                           # HMRM_ERR_INVALID


# ---------------------------------------------------------------------------------------------------------------
# device-side intermediates against the reference's own functions (frames are blind to ulp-level errors)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("traversal", [1, 2, 3, 4])
def test_device_rays_match_reference_kat(hmrm, renderer, oracle, traversal):
    """ImagePlane::GetRay on the DEVICE (src/Perspective.cpp:25-32, src/Spherical.cpp:17-31,
    src/Orthographic.cpp:19-25): bit patterns of ray pos / dir at the reference's known-answer pixels."""
    recs = H.golden_kat()["rays"]
    groups = {}
    for rec in recs:
        key = (rec["projection"], tuple(rec["pos"]), rec["hang"], rec["vang"], rec["hfov"], rec["ortho_width"], rec["W"], rec["H"])
        groups.setdefault(key, []).append(rec)
    scene = S.SCENE_BY_NAME["persp_basic"]
    H.configure(renderer, scene, H.load_scene_maps(scene, oracle))
    checked = 0
    for (proj, pos, hang, vang, hfov, ow, W, Hh), rs in groups.items():
        f = renderer.frame(projection=proj, screen_width=W, screen_height=Hh, cam_pos=[H.fx(v) for v in pos],
                           hang=H.fx(hang), vang=H.fx(vang), hfov=H.fx(hfov), ortho_width=H.fx(ow), grid_width=0.01,
                           step_dist=0.05, traversal=traversal, flags=hmrm.FLAG_RAY_DUMP)
        renderer.render(f)
        dump = renderer.ray_dump(f)
        for rec in rs:
            d = dump[rec["py"], rec["px"]]
            for i in range(3):
                assert H.same(d[i], rec["ray_pos"][i]), (proj, rec["px"], rec["py"], "pos", i)
                assert H.same(d[3 + i], rec["ray_dir"][i]), (proj, rec["px"], rec["py"], "dir", i)
            checked += 1
    assert checked == len(recs) and checked >= 300


def test_device_aabb_matches_reference_kat(hmrm, renderer):
    """distance() / intersection() on the DEVICE (src/AABB.cpp:30-77) on the reference's known answers: axis-parallel,
    grazing, inside-box, behind-box and zero-component rays (inf / NaN paths)."""
    recs = H.golden_kat()["aabb"]
    rays = np.array([[H.fx(v) for v in rec["pos"]] + [H.fx(v) for v in rec["dir"]] for rec in recs])
    boxes = np.array([[H.fx(v) for v in rec["c0"]] + [H.fx(v) for v in rec["c1"]] for rec in recs])
    out = renderer.debug_aabb(rays, boxes)
    for rec, o in zip(recs, out):
        assert H.same(o[0], rec["distance"]), rec
        assert int(o[1]) == rec["hit"], rec
        if rec["hit"]:
            for i in range(3):
                assert H.same(o[2 + i], rec["point"][i]), rec
    assert len(recs) >= 50


@pytest.mark.parametrize("name", ["persp_graze", "spher_wide", "ortho_fine", "camera_inside_box", "from_below", "defaults_grid"])
def test_device_ray_dump_matches_oracle_everywhere(hmrm, renderer, oracle, name):
    """Every pixel of a frame: device ray, slab distance and entry point == the oracle's (pinned to the reference)."""
    scene = S.SCENE_BY_NAME[name]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    f = H.product_frame(hmrm, renderer, scene, flags=hmrm.FLAG_RAY_DUMP)
    renderer.render(f)
    dump = renderer.ray_dump(f)
    of = H.oracle_frame(oracle, scene)
    W, Hh = scene["width"], scene["height"]
    gw = scene["grid_width"]
    mh, mw = maps[0].shape[:2]
    c0 = (0.0, 0.0, scene["min_height"])
    c1 = (c0[0] + mw * gw, c0[1] - mh * gw, scene["max_height"])
    rng = np.random.RandomState(3)
    pixels = [(0, 0), (W - 1, 0), (0, Hh - 1), (W - 1, Hh - 1)] + [(int(rng.randint(W)), int(rng.randint(Hh))) for _ in range(400)]
    entered = 0
    for px, py in pixels:
        pos, d = oracle.get_ray(of, px / (W - 1), py / (Hh - 1))
        rec = dump[py, px]
        assert all(H.bits(rec[i]) == H.bits(pos[i]) for i in range(3)), (px, py, "pos")
        assert all(H.bits(rec[3 + i]) == H.bits(d[i]) for i in range(3)), (px, py, "dir")
        assert H.bits(rec[6]) == H.bits(oracle.distance(pos, d, c0, c1)), (px, py, "distance")
        hit, point = oracle.intersection(pos, d, c0, c1)
        if hit:
            entered += 1
            assert all(H.bits(rec[7 + i]) == H.bits(point[i]) for i in range(3)), (px, py, "entry")
        else:
            assert not rec[7:].any()
    if name not in ("camera_inside_box",):
        assert entered > 0


# ---------------------------------------------------------------------------------------------------------------
# oracle parity at the BASELINE map sizes: sampled 4-row bands rendered by the oracle over the procedural map
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("workload,frame_no", [("sample720", 0), ("spherical1080", 31), ("ortho4k", 3),
                                                ("flythrough4k", 7), ("bands8k", 100)])
def test_baseline_full_size_bands_match_oracle(hmrm, oracle, workload, frame_no):
    """All five BASELINE configs at their full map sizes (1024^2 .. 32768^2) and resolutions: for sampled 4-row bands
    (top, around the horizon / first terrain rows, middle, bottom) the production kernel's pixels and first-hit step
    indices equal the oracle's.  The oracle generates the procedural map's texels on the fly (oracle_render_synth), so
    no 8 GiB height array is needed, and nothing the GPU computed feeds the expected values."""
    bench = load_bench()
    wl = bench.WORKLOADS[workload]
    W, Hh = wl["W"], wl["H"]
    c = bench.camera(wl, frame_no)
    r = hmrm.Renderer(0)
    try:
        r.min_height, r.max_height = bench.MIN_HEIGHT, bench.MAX_HEIGHT
        r.synth_maps(wl["log2n"], bench.SEED)
        f = r.frame(projection=wl["projection"], screen_width=W, screen_height=Hh, cam_pos=c["pos"],
                    hang=hmrm.deg2rad(c["hang_deg"]), vang=hmrm.deg2rad(c["vang_deg"]), hfov=hmrm.deg2rad(c["hfov_deg"]),
                    ortho_width=c["ortho_width"], grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"],
                    flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
        got = r.render(f).copy()
        steps = r.step_index(f).copy()
        assert r.stats().status == 0
    finally:
        r.close()
    hit_rows = np.nonzero((steps >= 0).any(axis=1))[0]
    assert hit_rows.size > 0
    first = int(hit_rows[0])
    starts = sorted({0, max(first - 2, 0), min(first + 40, Hh - 4), Hh // 2, (3 * Hh) // 4, Hh - 4})
    of = oracle.make_frame(projection=wl["projection"], width=W, height=Hh, grid_width=bench.GRID_WIDTH,
                           step_dist=wl["step_dist"], min_height=bench.MIN_HEIGHT, max_height=bench.MAX_HEIGHT, **c)
    total_hits = 0
    for a in starts:
        ofb, osteps, ost = oracle.render_synth(of, wl["log2n"], bench.SEED, rows=(a, a + 4))
        assert np.array_equal(got[a:a + 4], ofb[a:a + 4]), f"{workload}: rows {a}..{a + 3} differ from the oracle"
        assert np.array_equal(steps[a:a + 4], osteps[a:a + 4]), f"{workload}: step indices of rows {a}..{a + 3}"
        total_hits += ost.surf_hits
    assert total_hits > 0


# ---------------------------------------------------------------------------------------------------------------
# in-flight bookkeeping
# ---------------------------------------------------------------------------------------------------------------
def test_progressive_frame_after_inflight_whole_frames_does_not_disturb_them(hmrm, renderer, oracle):
    """Four whole frames through hmrm_render_async with no wait in between (all four device buffers in use, copies
    pending), then a progressive frame: it is rendered into buffer 0 on top of the newest picture, and must not
    overwrite buffer 0 while the first frame is still being copied out of it."""
    from heightmap_ray_marcher_b200 import binding

    scene = dict(S.SCENE_BY_NAME["persp_graze"], width=1600, height=900)
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    cams = [(-0.5 + 0.1 * i, 0.5 - 0.05 * i, 2.0 + 0.1 * i) for i in range(4)]
    want = [renderer.render(H.product_frame(hmrm, renderer, scene, cam_pos=c)).copy() for c in cams]
    for attempt in range(3):
        bufs = [binding.pinned_empty(want[0].shape) for _ in range(5)]
        for b in bufs:
            b[:] = 0
        for i, c in enumerate(cams):
            renderer.render_async(H.product_frame(hmrm, renderer, scene, cam_pos=c), bufs[i])
        renderer.render_async(H.product_frame(hmrm, renderer, scene, cam_pos=cams[0], cycle=2, cycle_period=3), bufs[4])
        renderer.wait()
        for i in range(4):
            assert np.array_equal(bufs[i], want[i]), f"attempt {attempt}: whole frame {i} was disturbed"
        sel = np.zeros(want[0].shape[0] * want[0].shape[1], dtype=bool)
        sel[2::3] = True
        sel = sel.reshape(want[0].shape[:2])
        assert np.array_equal(bufs[4][sel], want[0][sel])


def test_step_index_of_frames_in_flight_on_different_streams(hmrm, renderer, oracle):
    """hmrm_render_device on caller streams with HMRM_FLAG_STEP_INDEX: each launch owns its step-index buffer, so a
    second frame in flight (of a different size) cannot free or overwrite the first one's."""
    import torch

    scene = S.SCENE_BY_NAME["persp_graze"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    _, osteps, _ = H.oracle_render_scene(oracle, scene, maps)
    big = dict(scene, width=1280, height=720)
    s1, s2 = torch.cuda.Stream(device=0), torch.cuda.Stream(device=0)
    d1 = torch.zeros((big["height"], big["width"], 4), dtype=torch.uint8, device="cuda:0")
    d2 = torch.zeros((scene["height"], scene["width"], 4), dtype=torch.uint8, device="cuda:0")
    f_big = H.product_frame(hmrm, renderer, big, flags=hmrm.FLAG_STEP_INDEX)
    f_small = H.product_frame(hmrm, renderer, scene, flags=hmrm.FLAG_STEP_INDEX)
    renderer.render_device(f_small, d2, s2.cuda_stream)
    renderer.render_device(f_big, d1, s1.cuda_stream)          # grows nothing of the first launch's slot
    renderer.render_device(f_small, d2, s2.cuda_stream)
    got = renderer.step_index(f_small)
    torch.cuda.synchronize()
    assert np.array_equal(got, osteps)
