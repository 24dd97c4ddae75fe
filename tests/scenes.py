"""Seeded parity scenes shared by the golden generator, the oracle tests and the GPU tests.

Each scene = maps + lum/min/max + one frame of render state in the reference's own
vocabulary (config grammar names; angles in degrees as the grammar takes them).
Maps are either the integer fBm of csrc/synth_fbm.h or seeded random RGB noise.
"""
from __future__ import annotations

import numpy as np

DEFAULT_LUM = (0.299, 0.587, 0.114)     # main/hmap.cpp:45-47


def _scene(name, projection, maps, **kw):
    s = dict(name=name, projection=projection, maps=maps, lum=DEFAULT_LUM, width=320, height=180,
             pos=(-0.6, 0.6, 3.2), hang_deg=-45.0, vang_deg=120.0, hfov_deg=90.0, grid_width=0.01,
             step_dist=0.05, ortho_width=0.012, min_height=0.0, max_height=2.5, bg=(0, 0, 0))
    s.update(kw)
    return s


SYNTH8 = dict(kind="synth", log2n=8, seed=1234)            # 256 x 256
SYNTH9 = dict(kind="synth", log2n=9, seed=77)              # 512 x 512
NOISE = dict(kind="noise", w=200, h=136, seed=5)           # non-square, non-grey RGB
CROP = dict(kind="synth_crop", log2n=8, seed=1234, w=250, h=130)

SCENES = [
    # the three projections over the same terrain
    _scene("persp_basic", 1, SYNTH8),
    _scene("spher_basic", 2, SYNTH8),
    _scene("ortho_basic", 3, SYNTH8),
    # larger map, lower camera, long grazing marches
    _scene("persp_graze", 1, SYNTH9, pos=(-0.5, 0.5, 2.0), vang_deg=100.0, max_height=1.5, width=400, height=225),
    _scene("spher_wide", 2, SYNTH9, pos=(2.5, -2.5, 4.0), hang_deg=30.0, vang_deg=150.0, hfov_deg=170.0,
           max_height=2.0, width=384, height=216),
    _scene("ortho_fine", 3, SYNTH9, pos=(-1.0, 1.0, 5.0), vang_deg=130.0, step_dist=0.00125, ortho_width=0.02,
           max_height=2.0, width=256, height=144),
    # step_dist below the cell size (BASELINE config 3 shape), perspective and spherical
    _scene("persp_fine", 1, SYNTH8, step_dist=0.00125, width=200, height=120),
    _scene("spher_fine", 2, SYNTH8, step_dist=0.0031, width=200, height=120, bg=(10, 20, 30)),
    # quirks
    _scene("min_height_twice", 1, SYNTH8, min_height=0.75, max_height=2.5),            # D-1
    _scene("min_height_negative", 3, SYNTH8, min_height=-1.25, max_height=1.5, pos=(0.3, -0.3, 4.0), vang_deg=155.0),
    _scene("alpha_zero_centre", 3, SYNTH8, pos=(1.28, -1.28, 6.0), vang_deg=179.0, ortho_width=0.001,
           step_dist=0.01, bg=(200, 30, 90)),                                           # D-10, :1020
    _scene("camera_inside_box", 1, SYNTH8, pos=(1.0, -1.0, 1.0)),                        # D-3
    _scene("looking_up_sky", 2, SYNTH8, vang_deg=60.0, bg=(40, 50, 60)),                 # sky + bg tint
    _scene("looking_down_miss", 1, SYNTH8, pos=(-4.0, 4.0, 3.0), vang_deg=160.0, hang_deg=100.0, bg=(7, 8, 9)),
    _scene("top_down_ortho", 3, SYNTH8, pos=(1.28, -1.28, 9.0), vang_deg=180.0, hang_deg=0.0, ortho_width=0.009),
    _scene("from_below", 1, SYNTH8, pos=(1.0, -1.0, -2.0), vang_deg=20.0),               # enters through the floor
    # non-default luminance weights on non-grey RGB noise, non-square map
    _scene("noise_lum", 1, NOISE, lum=(0.9, 0.05, 0.3), pos=(-0.4, 0.5, 2.2), max_height=1.0, grid_width=0.013,
           step_dist=0.02),
    _scene("noise_lum_neg", 2, NOISE, lum=(-0.2, 1.4, 0.1), pos=(-0.4, 0.5, 2.2), max_height=1.0,
           grid_width=0.013, step_dist=0.02),
    _scene("crop_ortho", 3, CROP, pos=(-0.5, 0.4, 3.0), vang_deg=125.0),
    # big steps, odd resolution, default grid_width/step_dist of the reference (0.05 / 0.25)
    _scene("defaults_grid", 1, SYNTH8, grid_width=0.05, step_dist=0.25, pos=(-5.0, 5.0, 12.0), vang_deg=115.0,
           max_height=10.0, width=333, height=187),
    _scene("tiny_res", 2, SYNTH8, width=9, height=5),
]

SCENE_BY_NAME = {s["name"]: s for s in SCENES}


def build_maps(spec, synth_fn):
    """-> (height RGB8 [h,w,3], colormap RGBA8 [h,w,4]).  synth_fn(log2n, seed) provides the fBm maps."""
    if spec["kind"] == "synth":
        return synth_fn(spec["log2n"], spec["seed"])
    if spec["kind"] == "synth_crop":
        hm, cm = synth_fn(spec["log2n"], spec["seed"])
        return (np.ascontiguousarray(hm[3:3 + spec["h"], 2:2 + spec["w"]]),
                np.ascontiguousarray(cm[3:3 + spec["h"], 2:2 + spec["w"]]))
    if spec["kind"] == "noise":
        rng = np.random.RandomState(spec["seed"])
        h, w = spec["h"], spec["w"]
        # smooth-ish random field so that rays do hit something other than needles
        base = rng.randint(0, 256, size=(h // 8 + 2, w // 8 + 2, 3)).astype(np.float64)
        up = np.kron(base, np.ones((8, 8, 1)))[:h, :w]
        hm = np.clip(up + rng.randint(-20, 21, size=(h, w, 3)), 0, 255).astype(np.uint8)
        cm = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
        cm[::7, ::5, 3] = 0
        return hm, cm
    raise ValueError(spec["kind"])


def frame_kwargs(scene):
    keys = ("width", "height", "pos", "hang_deg", "vang_deg", "hfov_deg", "grid_width", "step_dist",
            "ortho_width", "min_height", "max_height", "bg")
    return {k: scene[k] for k in keys}
