"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the development container (where /root/reference is mounted):

    make -C oracle                 # builds oracle/_ref/{hmap_ref,ref_harness} from /root/reference
    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/frames.json   per scene: sha256 of the reference's RGBA8 framebuffer, step statistics
    tests/golden/frames.npz    the frames themselves (compressed), for diffing on the GPU box
    tests/golden/kat.json      known-answer vectors as C99 hex floats: DegreesToRads, GetRay for the three
                               ImagePlane classes, AABB distance()/intersection(), UpdateHeightmap samples

The reference ships no tests or vectors of its own, so these files are the pin:
frames come from oracle/_ref/hmap_ref (reference main loop under the fake SDL backend), vectors
from oracle/_ref/ref_harness (reference functions called directly).  The step statistics are the
oracle's (the reference does not count), recorded only after the oracle's frame matched bit for bit.
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

import oracle_lib as O  # noqa: E402
import scenes as S  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hx(v: float) -> str:
    return float(v).hex()


def golden_frames():
    meta, frames = {}, {}
    with tempfile.TemporaryDirectory(prefix="hmrm_gold_") as td:
        for sc in S.SCENES:
            hm, cm = S.build_maps(sc["maps"], O.synth_maps)
            hp, cp = f"{td}/{sc['name']}_h.png", f"{td}/{sc['name']}_c.png"
            O.write_png(hp, hm)
            O.write_png(cp, cm)
            kw = S.frame_kwargs(sc)
            cfg = O.config_text(kw, hp, cp, lum=sc["lum"])
            ref_frames, _ = O.run_ref(cfg, sc["projection"], kw["width"], kw["height"])
            ref = ref_frames[0]

            heights = O.update_heightmap(hm, sc["lum"], kw["min_height"], kw["max_height"])
            fr = O.make_frame(projection=sc["projection"], **kw)
            fb, steps, st = O.render(fr, heights, cm)
            if not np.array_equal(fb, ref):
                raise SystemExit(f"oracle != reference on scene {sc['name']}: "
                                 f"{int((fb != ref).any(axis=2).sum())} pixels differ")
            meta[sc["name"]] = dict(
                sha256=sha(ref), width=kw["width"], height=kw["height"], rays=st.rays, box_hits=st.box_hits,
                surf_hits=st.surf_hits, steps=st.steps, max_steps=st.max_steps, step_index_sha256=sha(steps),
                heights_sha256=sha(heights))
            frames[sc["name"]] = ref
            print(f"{sc['name']:22s} ok  box {st.box_hits:6d} surf {st.surf_hits:6d} steps {st.steps:9d} "
                  f"max {st.max_steps}")
    (HERE / "frames.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")
    np.savez_compressed(HERE / "frames.npz", **frames)


def golden_kat():
    rng = random.Random(20251018)
    kat = {"deg2rad": [], "rays": [], "aabb": [], "heights": []}

    # DegreesToRads
    degs = [0.0, 90.0, -45.0, 115.0, 125.0, 110.0, 180.0, 360.0, 1e-3, -719.25] + \
           [rng.uniform(-720, 720) for _ in range(20)]
    ans = O.ref_kat([f"G {hx(d)}" for d in degs])
    for d, a in zip(degs, ans):
        kat["deg2rad"].append({"deg": hx(d), "rad": a.split()[1]})

    # GetRay: three projections x several cameras x corner + random pixels
    cams = [
        dict(pos=(-5.0, 5.0, 0.0), hang=O.deg2rad(-45), vang=O.deg2rad(90), hfov=O.deg2rad(90), ow=0.1, W=1280, H=720),
        dict(pos=(-8.0, 8.0, 16.0), hang=O.deg2rad(-45), vang=O.deg2rad(115), hfov=O.deg2rad(90), ow=0.03, W=1920, H=1080),
        dict(pos=(-16.0, 16.0, 30.0), hang=O.deg2rad(-45), vang=O.deg2rad(125), hfov=O.deg2rad(60), ow=0.03, W=3840, H=2160),
        dict(pos=(221.92, -81.92, 40.0), hang=O.deg2rad(180), vang=O.deg2rad(110), hfov=O.deg2rad(90), ow=0.05, W=3840, H=2160),
        dict(pos=(1.5, -2.25, 7.0), hang=1.234567, vang=2.5, hfov=2.9, ow=0.0123, W=333, H=187),
        dict(pos=(0.1, 0.2, 0.3), hang=-3.0, vang=0.0, hfov=0.4, ow=1.0, W=9, H=5),
        dict(pos=(3.0, -3.0, 20.0), hang=0.5, vang=O.deg2rad(180), hfov=1.0, ow=0.02, W=640, H=480),
    ]
    queries, recs = [], []
    for cam in cams:
        pix = [(0, 0), (cam["W"] - 1, 0), (0, cam["H"] - 1), (cam["W"] - 1, cam["H"] - 1)] + \
              [(rng.randrange(cam["W"]), rng.randrange(cam["H"])) for _ in range(12)]
        for proj in (1, 2, 3):
            for (px, py) in pix:
                w, h = px / (cam["W"] - 1), py / (cam["H"] - 1)
                queries.append("R %d %s %s %s %s %s %s %s %d %d %s %s" % (
                    proj, hx(cam["pos"][0]), hx(cam["pos"][1]), hx(cam["pos"][2]), hx(cam["hang"]), hx(cam["vang"]),
                    hx(cam["hfov"]), hx(cam["ow"]), cam["W"], cam["H"], hx(w), hx(h)))
                recs.append(dict(projection=proj, pos=[hx(v) for v in cam["pos"]], hang=hx(cam["hang"]),
                                 vang=hx(cam["vang"]), hfov=hx(cam["hfov"]), ortho_width=hx(cam["ow"]),
                                 W=cam["W"], H=cam["H"], px=px, py=py, w=hx(w), h=hx(h)))
    for rec, a in zip(recs, O.ref_kat(queries)):
        t = a.split()
        rec["ray_pos"], rec["ray_dir"] = t[1:4], t[4:7]
        kat["rays"].append(rec)

    # AABB: generic, axis-parallel (zero components -> inf/NaN paths), grazing, inside, behind, negative-y box
    c0, c1 = (0.0, 0.0, 0.0), (10.24, -10.24, 10.0)
    cases = [
        ((-5, 5, 12), (0.5, -0.5, -0.70710678118654757), c0, c1),
        ((5, -5, 20), (0.0, 0.0, -1.0), c0, c1),                  # straight down: two zero components
        ((5, -5, 20), (0.0, 0.0, 1.0), c0, c1),                   # away
        ((5, -5, 5), (1.0, 0.0, 0.0), c0, c1),                    # origin inside the box
        ((-1, -5, 5), (1.0, 0.0, 0.0), c0, c1),                   # axis-parallel, enters x face
        ((-1, 1, 5), (1.0, 0.0, 0.0), c0, c1),                    # axis-parallel, misses in y
        ((0.0, -5, 5), (0.0, 1.0, 0.0), c0, c1),                  # origin on a slab plane with zero dir -> NaN
        ((-3, 0.0, 10.0), (1.0, -1e-300, -1e-300), c0, c1),       # grazing the top edge
        ((20, -20, 30), (-0.6, 0.6, -0.52915026221291817), c0, c1),
        ((-2, 2, 1), (0.70710678118654757, -0.70710678118654757, 0.0), (0.0, 0.0, 0.75), (2.56, -2.56, 2.5)),
        ((1.0, -1.0, -2.0), (0.1, -0.2, 0.97467943448089633), (0.0, 0.0, 0.0), (2.56, -2.56, 2.5)),
    ]
    for _ in range(40):
        o = (rng.uniform(-20, 30), rng.uniform(-30, 20), rng.uniform(-5, 40))
        d = [rng.uniform(-1, 1) for _ in range(3)]
        n = sum(v * v for v in d) ** 0.5
        cases.append((o, tuple(v / n for v in d), c0, c1))
    queries = ["D " + " ".join(hx(v) for grp in case for v in grp) for case in cases]
    for case, a in zip(cases, O.ref_kat(queries)):
        t = a.split()
        kat["aabb"].append(dict(pos=[hx(v) for v in case[0]], dir=[hx(v) for v in case[1]],
                                c0=[hx(v) for v in case[2]], c1=[hx(v) for v in case[3]],
                                distance=t[1], hit=int(t[2]), point=t[3:6]))

    # UpdateHeightmap on the full 8-bit cube corners + random RGB, several lum/min/max sets
    rgb_rng = np.random.RandomState(99)
    rgb = rgb_rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8)
    rgb[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 253], [128, 128, 128]]
    params = [
        ((0.299, 0.587, 0.114), 0.0, 10.0),
        ((0.299, 0.587, 0.114), 0.75, 2.5),
        ((1.0, 1.0, 1.0), -3.0, 4.5),           # clamps at 255
        ((-0.5, 0.25, 0.125), 1.0, 0.0),        # clamps at 0, inverted span
        ((1 / 3, 1 / 3, 1 / 3), 0.0, 1e-3),
    ]
    for lum, mn, mx in params:
        h = O.ref_heights(rgb, lum, mn, mx)
        kat["heights"].append(dict(lum=[hx(v) for v in lum], min_height=hx(mn), max_height=hx(mx),
                                   rgb_seed=99, sha256=sha(h), first_row=[hx(v) for v in h[0, :16]]))
    (HERE / "kat.json").write_text(json.dumps(kat, indent=1) + "\n")
    print(f"kat: {len(kat['deg2rad'])} deg2rad, {len(kat['rays'])} rays, {len(kat['aabb'])} aabb, "
          f"{len(kat['heights'])} height sets")


if __name__ == "__main__":
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle` first")
    golden_frames()
    golden_kat()
