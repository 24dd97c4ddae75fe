"""GPU: randomised / adversarial parity.  Small frames over small maps so that the CPU oracle can check every pixel:
framebuffer, per-pixel first-hit step index and step totals of BOTH traversals must equal the oracle's.

The parameter distributions are chosen to hit the corners of the skip traversal's exactness argument:
positions that cross many binades (rays leaving through the x = 0 / y = 0 edges, cameras far away or very close),
steps much smaller / much larger than a cell, grid widths that are powers of two (samples land exactly on cell
edges: the fixed-point ambiguity path), axis-parallel and exactly diagonal directions, negative and inverted height
ranges, maps that are tiny, non-square or flat."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def random_maps(rng, kind):
    if kind == "tiny":
        h, w = int(rng.randint(1, 6)), int(rng.randint(1, 6))
    elif kind == "strip":
        h, w = int(rng.randint(1, 4)), int(rng.randint(40, 300))
    else:
        h, w = int(rng.randint(20, 200)), int(rng.randint(20, 200))
    style = rng.randint(0, 4)
    if style == 0:       # white noise: needles; half of the cells grey (R == G == B: K1's table path), half coloured
        hm = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        grey = rng.rand(h, w) < 0.5
        hm[grey] = hm[grey][:, :1]
    elif style == 1:     # flat
        hm = np.full((h, w, 3), int(rng.randint(0, 256)), dtype=np.uint8)
    elif style == 2:     # smooth ramp + steps
        yy, xx = np.mgrid[0:h, 0:w]
        v = (xx * 255 // max(w - 1, 1) + (yy // 7) * 31) % 256
        hm = np.dstack([v, v, v]).astype(np.uint8)
    else:                # blocks
        base = rng.randint(0, 256, size=(h // 6 + 1, w // 6 + 1, 3))
        hm = np.kron(base, np.ones((6, 6, 1)))[:h, :w].astype(np.uint8)
    cm = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    cm[rng.rand(h, w) < 0.1, 3] = 0
    return hm, cm


def random_case(rng, i):
    kind = ["normal", "normal", "normal", "tiny", "strip"][i % 5]
    hm, cm = random_maps(rng, kind)
    h, w = hm.shape[:2]
    gw = float(rng.choice([0.01, 0.05, 0.013, 0.25, 0.0078125, 1.0, 3.7]))
    step_cells = float(rng.choice([0.05, 0.3, 1.0, 2.0, 5.0, 17.3, 400.0]))
    ext_x, ext_y = w * gw, h * gw
    mn, mx = [(0.0, 10.0), (0.0, 1.0), (-3.0, 2.0), (2.0, 0.5), (0.75, 2.5), (0.0, 0.0)][rng.randint(0, 6)]
    scale = max(ext_x, ext_y, abs(mx), abs(mn), gw)
    mode = rng.randint(0, 6)
    if mode == 0:      # far away
        pos = (ext_x / 2 + rng.uniform(-1, 1) * 300 * scale, -ext_y / 2 + rng.uniform(-1, 1) * 300 * scale, rng.uniform(50, 400) * scale)
    elif mode == 1:    # hugging the box
        pos = (rng.uniform(-0.1, 1.1) * ext_x, -rng.uniform(-0.1, 1.1) * ext_y, max(mn, mx) + rng.uniform(1e-6, 0.3) * scale)
    elif mode == 2:    # below / beside
        pos = (rng.uniform(-2, 3) * ext_x, -rng.uniform(-2, 3) * ext_y, rng.uniform(-2, 0.5) * scale)
    else:
        pos = (rng.uniform(-1.5, 2.5) * ext_x, -rng.uniform(-1.5, 2.5) * ext_y, max(mn, mx) + rng.uniform(0.05, 3) * scale)
    # look at a random point of the terrain box (so that most frames actually march), with some jitter
    tx, ty = rng.uniform(0.05, 0.95) * ext_x, -rng.uniform(0.05, 0.95) * ext_y
    tz = min(mn, mx) + rng.uniform(0.0, 1.0) * abs(mx - mn)
    dxy = np.hypot(tx - pos[0], ty - pos[1])
    dist = np.sqrt(dxy * dxy + (tz - pos[2]) ** 2) + 1e-30
    hang = float(np.arctan2(ty - pos[1], tx - pos[0])) + rng.uniform(-0.05, 0.05)
    vang = float(np.arccos(np.clip((tz - pos[2]) / dist, -1.0, 1.0))) + rng.uniform(-0.05, 0.05)
    vang = float(np.clip(vang, 0.0, np.pi))
    if rng.rand() < 0.12:
        hang = float(rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, np.pi / 4]))     # axis-parallel / diagonal
    if rng.rand() < 0.08:
        vang = float(rng.choice([np.pi, np.pi / 2, 3.0]))
    fov_to_box = 2.0 * np.arctan2(0.7 * max(ext_x, ext_y), dist)                       # the box about fills the frame
    return dict(hm=hm, cm=cm, lum=[(0.299, 0.587, 0.114), (1.0, 0.0, 0.0), (0.2, -0.1, 0.9)][rng.randint(0, 3)],
                min_height=mn, max_height=mx,
                frame=dict(projection=int(rng.randint(1, 4)), screen_width=int(rng.randint(2, 70)),
                           screen_height=int(rng.randint(2, 40)), cam_pos=pos, hang=hang, vang=vang,
                           hfov=float(np.clip(fov_to_box * rng.uniform(0.5, 2.0), 0.02, 3.0)),
                           ortho_width=float(rng.uniform(0.3, 2.0) * max(ext_x, ext_y) / 50),
                           grid_width=gw, step_dist=gw * step_cells,
                           bg=(int(rng.randint(0, 256)), int(rng.randint(0, 256)), int(rng.randint(0, 256)))))


@pytest.mark.parametrize("seed", range(48))
def test_random_cases_match_oracle(hmrm, renderer, oracle, seed):
    rng = np.random.RandomState(4242 + seed)
    with_hits = with_box = long_marches = 0
    for i in range(25):
        case = random_case(rng, i)
        renderer.lum_r, renderer.lum_g, renderer.lum_b = case["lum"]
        renderer.min_height, renderer.max_height = case["min_height"], case["max_height"]
        renderer.set_maps(case["hm"], case["cm"])
        fk = case["frame"]
        of = oracle.make_frame(projection=fk["projection"], width=fk["screen_width"], height=fk["screen_height"],
                               pos=fk["cam_pos"], hang=fk["hang"], vang=fk["vang"], hfov=fk["hfov"],
                               ortho_width=fk["ortho_width"], grid_width=fk["grid_width"], step_dist=fk["step_dist"],
                               min_height=case["min_height"], max_height=case["max_height"], bg=fk["bg"])
        heights = oracle.update_heightmap(case["hm"], case["lum"], case["min_height"], case["max_height"])
        # the reference never terminates for some rays; the oracle caps at 2^31 steps: skip those cases
        # by bounding the work first with the brute kernel's own cut-off flag
        fb = renderer.frame(traversal=1, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX, **fk)
        a = renderer.render(fb).copy()
        sa, ia = renderer.stats(), renderer.step_index(fb)
        tag = f"seed {seed} case {i}: {fk} map {case['hm'].shape} lum {case['lum']} h [{case['min_height']},{case['max_height']}]"
        for trav in (2, 3, 4):
            fs = renderer.frame(traversal=trav, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX, **fk)
            b = renderer.render(fs).copy()
            sb, ib = renderer.stats(), renderer.step_index(fs)
            assert sa.status == sb.status, (trav, tag)
            assert np.array_equal(a, b), (trav, tag)
            assert np.array_equal(ia, ib), (trav, tag)
            assert (sa.steps, sa.box_hits, sa.surf_hits, sa.max_steps) == (sb.steps, sb.box_hits, sb.surf_hits, sb.max_steps), (trav, tag)
        if sa.status == 0 and sa.max_steps < 3_000_000:
            want, osteps, ost = oracle.render(of, heights, case["cm"])
            assert np.array_equal(a, want), tag
            assert np.array_equal(ia, osteps), tag
            assert (sa.steps, sa.box_hits, sa.surf_hits) == (ost.steps, ost.box_hits, ost.surf_hits), tag
        with_hits += sa.surf_hits > 0
        with_box += sa.box_hits > 0
        long_marches += sa.max_steps > 200
    # the fuzz must not be vacuous: most cases enter the box, many hit terrain, some march far
    assert with_box >= 8 and with_hits >= 5, (with_box, with_hits, long_marches)


def check_against_oracle(hmrm, renderer, oracle, hm, cm, lum, mn, mx, fk, tag):
    renderer.lum_r, renderer.lum_g, renderer.lum_b = lum
    renderer.min_height, renderer.max_height = mn, mx
    renderer.set_maps(hm, cm)
    of = oracle.make_frame(projection=fk["projection"], width=fk["screen_width"], height=fk["screen_height"],
                           pos=fk["cam_pos"], hang=fk["hang"], vang=fk["vang"], hfov=fk["hfov"],
                           ortho_width=fk["ortho_width"], grid_width=fk["grid_width"], step_dist=fk["step_dist"],
                           min_height=mn, max_height=mx, bg=fk["bg"])
    heights = oracle.update_heightmap(hm, lum, mn, mx)
    want, osteps, ost = oracle.render(of, heights, cm)
    out = []
    for trav in (1, 2, 3, 4):
        f = renderer.frame(traversal=trav, flags=hmrm.FLAG_STATS | hmrm.FLAG_STEP_INDEX, **fk)
        got = renderer.render(f).copy()
        st, si = renderer.stats(), renderer.step_index(f)
        assert np.array_equal(got, want), f"{tag} traversal {trav}: {int((got != want).any(axis=2).sum())} pixels differ"
        assert np.array_equal(si, osteps), f"{tag} traversal {trav}: step indices differ"
        assert (st.steps, st.box_hits, st.surf_hits, st.max_steps) == (ost.steps, ost.box_hits, ost.surf_hits, ost.max_steps), tag
        out.append(st)
    return out


@pytest.mark.parametrize("seed", range(24))
def test_long_grazing_marches_match_oracle(hmrm, renderer, oracle, seed):
    """Thousands of steps per ray at low clearance over rough and over flat terrain: long chains of jumps, many
    binade crossings (coordinates run from ~100 cells down through 1, 0.5, 0.25 ... towards the x = 0 / y = 0 edges)."""
    rng = np.random.RandomState(9000 + seed)
    n = int(rng.choice([300, 512, 700]))
    yy, xx = np.mgrid[0:n, 0:n]
    style = seed % 4
    if style == 0:
        v = (np.sin(xx / 37.0) * np.cos(yy / 23.0) * 90 + 110 + rng.randint(0, 25, size=(n, n))).clip(0, 255)
    elif style == 1:
        v = np.full((n, n), 40.0)
        v[n // 3: n // 3 + 9, :] = 200      # a wall across a plain
    elif style == 2:
        v = rng.randint(0, 256, size=(n // 16 + 1, n // 16 + 1))
        v = np.kron(v, np.ones((16, 16)))[:n, :n]
    else:
        v = ((xx + yy) * 255.0 / (2 * n - 2))
    hm = np.dstack([v, v, v]).astype(np.uint8)
    cm = rng.randint(0, 256, size=(n, n, 4)).astype(np.uint8)
    gw = float(rng.choice([0.01, 0.03125, 0.2]))
    ext = n * gw
    mx = float(rng.choice([0.05, 0.3, 1.0])) * ext
    side = rng.randint(0, 4)
    along = rng.uniform(0.1, 0.9) * ext
    off = rng.uniform(0.02, 0.4) * ext
    pos = [(-off, -along), (ext + off, -along), (along, off), (along, -ext - off)][side]
    height = mx * rng.uniform(0.3, 1.6)
    target = (rng.uniform(0.2, 0.8) * ext, -rng.uniform(0.2, 0.8) * ext, rng.uniform(0.0, 0.6) * mx)
    d = np.array([target[0] - pos[0], target[1] - pos[1], target[2] - height])
    fk = dict(projection=int(rng.randint(1, 4)), screen_width=int(rng.randint(24, 64)), screen_height=int(rng.randint(12, 36)),
              cam_pos=(pos[0], pos[1], height), hang=float(np.arctan2(d[1], d[0])),
              vang=float(np.arccos(d[2] / np.linalg.norm(d))), hfov=float(rng.uniform(0.3, 1.4)),
              ortho_width=float(ext * rng.uniform(0.3, 1.0) / 64), grid_width=gw,
              step_dist=gw * float(rng.choice([0.07, 0.25, 0.6, 1.0])), bg=(1, 2, 3))
    st = check_against_oracle(hmrm, renderer, oracle, hm, cm, (0.299, 0.587, 0.114), 0.0, mx, fk, f"seed {seed} {fk}")
    assert st[0].max_steps > 100, st[0].max_steps
    assert st[1].fetches < st[0].fetches          # the skip traversal really skipped


@pytest.mark.parametrize("case", range(6))
def test_rays_longer_than_the_model_period_match_oracle(hmrm, renderer, oracle, case):
    """Steps of 1/4000 .. 1/700 of a cell: hundreds of thousands of samples per ray, so the integer model of
    k2_render_lin is re-anchored on an exact position several times per ray (HMRM_LIN_PERIOD = 65536 samples) and the
    FP64 skip kernel crosses many binades.  A handful of rays, every one checked against the oracle."""
    rng = np.random.RandomState(7700 + case)
    n = [96, 128, 160, 64, 200, 100][case]
    yy, xx = np.mgrid[0:n, 0:n]
    if case % 3 == 0:
        v = np.full((n, n), 30.0)
        v[:, n // 2: n // 2 + 3] = 220                       # a wall far from the entry edge
    elif case % 3 == 1:
        v = (np.sin(xx / 11.0) * np.cos(yy / 17.0) * 60 + 90).clip(0, 255)
    else:
        v = (xx * 200.0 / (n - 1))
    hm = np.dstack([v, v, v]).astype(np.uint8)
    cm = rng.randint(0, 256, size=(n, n, 4)).astype(np.uint8)
    gw = [0.01, 0.05, 0.0078125, 1.0, 0.02, 0.3][case]
    ext = n * gw
    mx = 0.25 * ext
    step = gw / [2000.0, 700.0, 1200.0, 4000.0, 1500.0, 2500.0][case]
    height = mx * 1.05
    pos = (-0.05 * ext, -0.45 * ext, height)
    target = (0.9 * ext, -0.55 * ext, 0.1 * mx)
    d = np.array([target[0] - pos[0], target[1] - pos[1], target[2] - pos[2]])
    fk = dict(projection=1 + case % 3, screen_width=4, screen_height=3, cam_pos=pos,
              hang=float(np.arctan2(d[1], d[0])), vang=float(np.arccos(d[2] / np.linalg.norm(d))), hfov=0.2,
              ortho_width=float(ext / 40), grid_width=gw, step_dist=step, bg=(9, 8, 7))
    st = check_against_oracle(hmrm, renderer, oracle, hm, cm, (0.299, 0.587, 0.114), 0.0, mx, fk, f"case {case} {fk}")
    assert st[0].max_steps > 66000, st[0].max_steps
