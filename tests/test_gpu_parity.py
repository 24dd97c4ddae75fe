"""GPU: the CUDA path, called through the C ABI (libhmrm.so), against the CPU oracle on the same seeded
inputs and against the committed fixtures produced by the unmodified reference.

Bar: bit-exact RGBA8 framebuffer, identical per-pixel first-hit step index, identical step statistics
(FP64 exact mode).  Full-size cases are checked through size-independent properties.
"""
import numpy as np
import pytest

import helpers as H
import scenes as S

pytestmark = pytest.mark.gpu

TRAVERSALS = [1, 2, 3, 4]   # HMRM_TRAVERSAL_BRUTE, _SKIP (integer linear model), _SKIP_FP64, _PACK (refilled lanes)


@pytest.mark.parametrize("traversal", TRAVERSALS)
@pytest.mark.parametrize("scene", S.SCENES, ids=lambda s: s["name"])
def test_frame_bit_exact_vs_reference_golden_and_oracle(hmrm, renderer, oracle, scene, traversal):
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    f = H.product_frame(hmrm, renderer, scene, traversal=traversal,
                        flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
    got = renderer.render(f)
    st = renderer.stats()
    steps = renderer.step_index(f)

    meta = H.golden_meta()[scene["name"]]
    want = H.golden_frames()[scene["name"]]
    diff = int((got != want).any(axis=2).sum())
    assert diff == 0, f"{diff} pixels differ from the reference's frame"
    assert H.sha(got) == meta["sha256"]

    ofb, osteps, ost = H.oracle_render_scene(oracle, scene, maps)
    assert np.array_equal(got, ofb)
    assert np.array_equal(steps, osteps), f"{int((steps != osteps).sum())} first-hit step indices differ"
    assert (st.rays, st.box_hits, st.surf_hits, st.steps, st.max_steps) == (
        ost.rays, ost.box_hits, ost.surf_hits, ost.steps, ost.max_steps)
    assert st.status == 0
    if traversal == 1:
        assert st.fetches == st.steps
    else:
        assert st.fetches <= st.steps + st.box_hits * 64


@pytest.mark.parametrize("name", ["noise_lum", "noise_lum_neg", "min_height_twice", "defaults_grid"])
def test_prepass_heights_bit_exact(hmrm, renderer, oracle, name):
    scene = S.SCENE_BY_NAME[name]
    hm, cm = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, (hm, cm))
    got = renderer.heights()
    want = oracle.update_heightmap(hm, scene["lum"], scene["min_height"], scene["max_height"])
    assert got.tobytes() == want.tobytes()
    assert H.sha(got) == H.golden_meta()[name]["heights_sha256"]


def test_prepass_heights_kat(hmrm, renderer):
    """UpdateHeightmap known answers from the unmodified reference (tests/golden/kat.json)."""
    for rec in H.golden_kat()["heights"]:
        rgb = np.random.RandomState(rec["rgb_seed"]).randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
        rgb[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1],
                      [254, 255, 253], [128, 128, 128]]
        renderer.lum_r, renderer.lum_g, renderer.lum_b = [H.fx(v) for v in rec["lum"]]
        renderer.min_height, renderer.max_height = H.fx(rec["min_height"]), H.fx(rec["max_height"])
        renderer.set_maps(rgb, np.zeros((64, 64, 4), dtype=np.uint8))
        assert H.sha(renderer.heights()) == rec["sha256"]


def test_synth_maps_device_equals_cpu_generator(hmrm, renderer, oracle):
    for log2n in (6, 9):
        renderer.synth_maps(log2n, 1234)
        hm, cm = renderer.get_maps()
        ohm, ocm = oracle.synth_maps(log2n, 1234)
        assert np.array_equal(hm, ohm) and np.array_equal(cm, ocm)


@pytest.mark.parametrize("traversal", TRAVERSALS)
def test_cycle_interleave_accumulates_to_full_frame(hmrm, renderer, oracle, traversal):
    """cycle / cycle_period (main/hmap.cpp:976-981): the persistent framebuffer is not cleared between
    frames, so cycle_period frames with successive phases add up to the full image."""
    scene = S.SCENE_BY_NAME["spher_basic"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    full = renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal)).copy()
    # different resolution forces a fresh (zeroed) framebuffer, then back
    renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal, screen_width=64, screen_height=32))
    period = 47   # the reference's default (main/hmap.cpp:71)
    out = None
    for phase in range(period):
        out = renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal, cycle=phase,
                                              cycle_period=period))
        if phase == 0:
            first = out.copy()
    assert np.array_equal(out, full)
    sel = np.zeros(full.shape[:2], dtype=bool).reshape(-1)
    sel[0::period] = True
    sel = sel.reshape(full.shape[:2])
    assert np.array_equal(first[sel], full[sel]) and not first[~sel].any()


@pytest.mark.parametrize("traversal", TRAVERSALS)
def test_row_bands_concatenate_to_full_frame(hmrm, renderer, oracle, traversal):
    """Row-band partition used for multi-GPU frames: bands rendered separately == one full frame."""
    scene = S.SCENE_BY_NAME["persp_graze"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    full = renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal)).copy()
    Hh = scene["height"]
    out = np.zeros_like(full)
    edges = [0, 57, 58, 113, 200, Hh]
    for a, b in zip(edges[:-1], edges[1:]):
        band = np.zeros_like(full)
        renderer.render(H.product_frame(hmrm, renderer, scene, traversal=traversal, row_begin=a, row_end=b), out=band)
        out[a:b] = band[a:b]
    assert np.array_equal(out, full)


def test_render_device_and_async_paths_agree(hmrm, renderer, oracle):
    import torch

    scene = S.SCENE_BY_NAME["ortho_basic"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    f = H.product_frame(hmrm, renderer, scene)
    want = renderer.render(f).copy()
    dev = torch.zeros((scene["height"], scene["width"], 4), dtype=torch.uint8, device="cuda:0")
    renderer.render_device(f, dev)
    renderer.wait()
    assert np.array_equal(dev.cpu().numpy(), want)
    from heightmap_ray_marcher_b200 import binding

    pinned = binding.pinned_empty(want.shape)
    pinned[:] = 0
    renderer.render_async(f, pinned)
    renderer.wait()
    assert np.array_equal(pinned, want)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_streaming_async_frames_equal_synchronous_frames(hmrm, renderer, oracle, depth):
    """hmrm_render_async + hmrm_wait_pending(depth): depth+1 frames in flight (copy-out overlaps the next kernels)."""
    from heightmap_ray_marcher_b200 import binding

    scene = S.SCENE_BY_NAME["persp_graze"]
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    cams = [(-0.5 + 0.1 * i, 0.5 - 0.05 * i, 2.0 + 0.1 * i) for i in range(8)]
    want = [renderer.render(H.product_frame(hmrm, renderer, scene, cam_pos=c)).copy() for c in cams]
    nb = depth + 1
    bufs = [binding.pinned_empty(want[0].shape) for _ in range(nb)]
    got = []
    for i, c in enumerate(cams):
        renderer.render_async(H.product_frame(hmrm, renderer, scene, cam_pos=c), bufs[i % nb])
        renderer.wait_pending(depth)
        if i >= depth:
            got.append(bufs[(i - depth) % nb].copy())  # frame i-depth is complete, the newer ones may be in flight
    renderer.wait()
    for i in range(len(cams) - depth, len(cams)):
        got.append(bufs[i % nb].copy())
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), f"frame {i}"
    # a progressive (cycle_period > 1) frame after whole frames lands on top of the newest picture
    f = H.product_frame(hmrm, renderer, scene, cam_pos=cams[0], cycle=3, cycle_period=5)
    out = renderer.render(f)
    sel = np.zeros(want[0].shape[0] * want[0].shape[1], dtype=bool)
    sel[3::5] = True
    sel = sel.reshape(want[0].shape[:2])
    assert np.array_equal(out[sel], want[0][sel]) and np.array_equal(out[~sel], want[-1][~sel])


@pytest.mark.parametrize("shrink", ["0.3", "0.02", "0"])
@pytest.mark.parametrize("name", ["persp_basic", "ortho_fine", "noise_lum_neg", "crop_ortho"])
def test_quantiser_range_never_affects_results(hmrm, oracle, name, shrink, monkeypatch):
    """The 16-bit height quantiser's range is only estimated from a sample of the map; values outside clamp on both
    sides of the comparison and tie (FP64 decides).  Force a far too narrow range: frames and step indices must
    still be bit-exact, only the number of fetches may change."""
    monkeypatch.setenv("HMRM_ZQ_RANGE_SHRINK", shrink)
    scene = S.SCENE_BY_NAME[name]
    maps = H.load_scene_maps(scene, oracle)
    r = hmrm.Renderer(0)
    try:
        H.configure(r, scene, maps)
        f = H.product_frame(hmrm, r, scene, traversal=2, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
        got = r.render(f)
        steps = r.step_index(f)
        st = r.stats()
    finally:
        r.close()
    ofb, osteps, ost = H.oracle_render_scene(oracle, scene, maps)
    assert np.array_equal(got, ofb) and np.array_equal(steps, osteps)
    assert (st.steps, st.surf_hits, st.box_hits) == (ost.steps, ost.surf_hits, ost.box_hits)


@pytest.mark.parametrize("scene", S.SCENES, ids=lambda s: s["name"])
def test_fp32_fast_mode_tolerance(hmrm, renderer, oracle, scene):
    """HMRM_FP32_FAST (north star: >= 99.5 % of pixels identical to the reference frame, every other pixel explained
    by a first-hit step index differing by at most 1).  The mode only uses FP32 as a filter with FP64 fallbacks, so
    the expected outcome is stronger: every pixel and every step index identical.  Tolerance written out anyway."""
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    f = H.product_frame(hmrm, renderer, scene, precision=hmrm.FP32_FAST, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
    got = renderer.render(f)
    steps = renderer.step_index(f)
    want = H.golden_frames()[scene["name"]]
    _, osteps, ost = H.oracle_render_scene(oracle, scene, maps)
    differs = (got != want).any(axis=2)
    assert 1.0 - differs.mean() >= 0.995
    assert (np.abs(steps[differs].astype(np.int64) - osteps[differs]) <= 1).all()
    assert not differs.any() and np.array_equal(steps, osteps)        # what the construction actually guarantees
    assert renderer.stats().steps == ost.steps


@pytest.mark.parametrize("proj", [1, 2, 3])
def test_fp32_fast_mode_equals_exact_on_random_cameras(hmrm, renderer, oracle, proj):
    """Randomised views (many of them mostly sky / mostly outside the map): FAST == EXACT bit for bit."""
    renderer.lum_r, renderer.lum_g, renderer.lum_b = S.DEFAULT_LUM
    renderer.min_height, renderer.max_height = 0.0, 10.0
    renderer.synth_maps(10, 99)
    rng = np.random.RandomState(1000 + proj)
    for i in range(24):
        kw = dict(projection=proj, screen_width=int(rng.randint(64, 700)), screen_height=int(rng.randint(48, 400)),
                  cam_pos=(rng.uniform(-15, 25), rng.uniform(-25, 15), rng.uniform(-5, 40)),
                  hang=rng.uniform(-3.2, 3.2), vang=rng.uniform(0.0, 3.14159), hfov=rng.uniform(0.3, 2.9),
                  grid_width=0.01, step_dist=float(rng.choice([0.05, 0.013, 0.2])), ortho_width=rng.uniform(0.005, 0.08),
                  bg=(int(rng.randint(0, 256)), int(rng.randint(0, 256)), int(rng.randint(0, 256))),
                  flags=hmrm.FLAG_STEP_INDEX)
        fe = renderer.frame(precision=hmrm.FP64_EXACT, **kw)
        a = renderer.render(fe).copy()
        ia = renderer.step_index(fe)
        ff = renderer.frame(precision=hmrm.FP32_FAST, **kw)
        b = renderer.render(ff).copy()
        ib = renderer.step_index(ff)
        assert np.array_equal(a, b) and np.array_equal(ia, ib), f"camera {i}: {int((a != b).any(axis=2).sum())} pixels"


def test_invalid_arguments_are_rejected(hmrm, renderer, oracle):
    scene = S.SCENE_BY_NAME["persp_basic"]
    H.configure(renderer, scene, H.load_scene_maps(scene, oracle))
    for bad in (dict(step_dist=0.0), dict(step_dist=-1.0), dict(grid_width=0.0), dict(cycle_period=0),
                dict(projection=4), dict(screen_width=1), dict(row_begin=10, row_end=5),
                dict(step_dist=float("nan"))):
        with pytest.raises(hmrm.HmrmError):
            renderer.render(H.product_frame(hmrm, renderer, scene, **bad))


def test_nonterminating_ray_is_cut_off_and_flagged(hmrm, renderer, oracle):
    """straight-up orthographic rays under a flat zero terrain never leave the grid in
    the reference (it hangs); the kernel must terminate and flag the frame."""
    hm = np.zeros((32, 32, 3), dtype=np.uint8)
    cm = np.full((32, 32, 4), 255, dtype=np.uint8)
    renderer.lum_r, renderer.lum_g, renderer.lum_b = 0.0, 0.0, 0.0
    renderer.min_height, renderer.max_height = 0.0, 10.0
    renderer.set_maps(hm, cm)
    for traversal in TRAVERSALS:
        f = renderer.frame(projection=3, screen_width=16, screen_height=8, cam_pos=(0.8, -0.8, -3.0),
                           hang=0.0, vang=0.0, ortho_width=0.01, grid_width=0.05, step_dist=0.25,
                           traversal=traversal, flags=hmrm.FLAG_STATS)
        out = renderer.render(f)
        st = renderer.stats()
        assert st.status == 4   # HMRM_ERR_NONTERMINATING
        assert (out[..., 3] == 255).all()


@pytest.mark.parametrize("log2n,proj,res", [(11, 1, (1920, 1080)), (12, 2, (1920, 1080))])
def test_large_frame_properties(hmrm, renderer, oracle, log2n, proj, res):
    """Sizes the oracle cannot cover quickly in full: BRUTE == SKIP (frame, step index, step totals),
    a sampled row band == oracle, alpha is 255 everywhere."""
    renderer.lum_r, renderer.lum_g, renderer.lum_b = S.DEFAULT_LUM
    renderer.min_height, renderer.max_height = 0.0, 10.0
    renderer.synth_maps(log2n, 1234)
    n = 1 << log2n
    W, Hh = res
    common = dict(projection=proj, screen_width=W, screen_height=Hh, cam_pos=(-0.2 * n * 0.01, 0.2 * n * 0.01, 16.0),
                  hang=hmrm.deg2rad(-45.0), vang=hmrm.deg2rad(115.0), hfov=hmrm.deg2rad(90.0), grid_width=0.01,
                  step_dist=0.05, ortho_width=0.03, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
    fb = renderer.frame(traversal=1, **common)
    a = renderer.render(fb).copy()
    sa, ia = renderer.stats(), renderer.step_index(fb)
    fs = renderer.frame(traversal=2, **common)
    b = renderer.render(fs).copy()
    sb, ib = renderer.stats(), renderer.step_index(fs)
    assert np.array_equal(a, b) and np.array_equal(ia, ib)
    assert (sa.rays, sa.box_hits, sa.surf_hits, sa.steps, sa.max_steps) == (sb.rays, sb.box_hits, sb.surf_hits,
                                                                             sb.steps, sb.max_steps)
    assert (a[..., 3] == 255).all() and sa.surf_hits > 0

    hm, cm = renderer.get_maps()
    heights = oracle.update_heightmap(hm, S.DEFAULT_LUM, 0.0, 10.0)
    of = oracle.make_frame(projection=proj, width=W, height=Hh, pos=common["cam_pos"], hang_deg=-45.0,
                           vang_deg=115.0, hfov_deg=90.0, grid_width=0.01, step_dist=0.05, ortho_width=0.03)
    rows = (Hh // 2 - 12, Hh // 2 + 12)
    ofb, osteps, _ = oracle.render(of, heights, cm, rows=rows)
    assert np.array_equal(a[rows[0]:rows[1]], ofb[rows[0]:rows[1]])
    assert np.array_equal(ia[rows[0]:rows[1]], osteps[rows[0]:rows[1]])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_interleaved_bands_compose_full_frame(hmrm, renderer, oracle, world):
    """Single-frame multi-GPU partition (BASELINE configs[4]) emulated on one GPU: every 'rank' renders its
    interleaved tile rows into its own buffer; packing + unpacking as the gather does gives the 1-GPU frame."""
    import torch

    from heightmap_ray_marcher_b200 import multi_gpu as MG

    scene = dict(S.SCENE_BY_NAME["persp_graze"], height=226)      # not a multiple of 4 * world
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    full = renderer.render(H.product_frame(hmrm, renderer, scene)).copy()
    Hh, W = scene["height"], scene["width"]
    T = MG.tile_rows(Hh)
    out = torch.zeros((T, 4, W, 4), dtype=torch.uint8, device="cuda:0")
    for r in range(world):
        local = torch.zeros((MG.padded_height(Hh), W, 4), dtype=torch.uint8, device="cuda:0")
        renderer.render_device(H.product_frame(hmrm, renderer, scene, band_count=world, band_index=r), local)
        renderer.wait()
        packed = MG.pack_owned(local, Hh, r, world)
        n = len(MG.owned_tile_rows(Hh, r, world))
        out[r::world] = packed[:n]
        # rows of other ranks were not touched
        mask = np.ones(MG.padded_height(Hh), dtype=bool)
        mask[MG.owned_pixel_rows(Hh, r, world)] = False
        assert not local.cpu().numpy()[mask].any()
    got = out.view(T * 4, W, 4)[:Hh].cpu().numpy()
    assert np.array_equal(got, full)


@pytest.mark.parametrize("workload,frame_no", [("flythrough4k", 7), ("spherical1080", 31), ("ortho4k", 3), ("bands8k", 100)])
def test_baseline_full_size_traversals_agree(hmrm, workload, frame_no):
    """BASELINE.json's full sizes (16384^2 at 4K, 8192^2 ortho with step_dist/8, 32768^2 at 8K), bench.py's own
    cameras: the reference loop as written (BRUTE) and both skip traversals give the same frame, the same per-pixel
    first-hit sample index and the same reference-equivalent step totals.  (bench.py pins the same workload against
    the unmodified reference on a 1/16 sample of the rays: cpu_baseline.parity_of_sample.)"""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("hmrm_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.WORKLOADS[workload]
    r = hmrm.Renderer(0)
    try:
        r.min_height, r.max_height = bench.MIN_HEIGHT, bench.MAX_HEIGHT
        r.synth_maps(wl["log2n"], bench.SEED)
        c = bench.camera(wl, frame_no)
        common = dict(projection=wl["projection"], screen_width=wl["W"], screen_height=wl["H"], cam_pos=c["pos"],
                      hang=hmrm.deg2rad(c["hang_deg"]), vang=hmrm.deg2rad(c["vang_deg"]), hfov=hmrm.deg2rad(c["hfov_deg"]),
                      ortho_width=c["ortho_width"], grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"],
                      flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX)
        got = []
        for trav in TRAVERSALS:
            f = r.frame(traversal=trav, **common)
            fb = r.render(f).copy()
            st = r.stats()
            got.append((fb, r.step_index(f).copy(), (st.rays, st.box_hits, st.surf_hits, st.steps, st.max_steps), st.status))
        base = got[0]
        assert base[3] == 0 and base[2][0] == wl["W"] * wl["H"] and base[2][2] > 0
        assert (base[0][..., 3] == 255).all()
        for other in got[1:]:
            assert np.array_equal(base[0], other[0])
            assert np.array_equal(base[1], other[1])
            assert base[2] == other[2] and other[3] == 0
    finally:
        r.close()


@pytest.mark.parametrize("world,height", [(2, 226), (3, 225), (8, 227), (5, 16)])
def test_bands_fill_one_host_frame(hmrm, renderer, oracle, world, height):
    """hmrm_render_async with band_count > 1 copies only this band's tile rows to the host (a strided copy plus the
    ragged last tile row), so several ranks can fill ONE host frame, each over its own PCIe link."""
    from heightmap_ray_marcher_b200 import binding

    scene = dict(S.SCENE_BY_NAME["persp_graze"], height=height)
    maps = H.load_scene_maps(scene, oracle)
    H.configure(renderer, scene, maps)
    full = renderer.render(H.product_frame(hmrm, renderer, scene)).copy()
    host = binding.pinned_empty(full.shape)
    host[:] = 0x5A
    for r in range(world):
        before = host.copy()
        renderer.render_async(H.product_frame(hmrm, renderer, scene, band_count=world, band_index=r), host)
        renderer.wait()
        own = np.zeros(height, dtype=bool)
        for t in range(r, (height + 3) // 4, world):
            own[t * 4: min(t * 4 + 4, height)] = True
        assert np.array_equal(host[own], full[own]), f"rank {r}: own rows"
        assert np.array_equal(host[~own], before[~own]), f"rank {r}: rows of other ranks were touched"
    assert np.array_equal(host, full)
