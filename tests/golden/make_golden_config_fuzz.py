"""Random config texts through the UNMODIFIED reference (oracle/_ref/hmap_ref, one frame against the fake SDL): what
ConsumeConfigStream (main/hmap.cpp:309-520) echoes, warns about and exits with for token soups — numbers in every
spelling iostreams accept or refuse, unknown identifiers, missing arguments, image paths that load, conflict or do
not exist.  tests/test_host_cli.py replays them through our `hmap --parse-only`.

    make -C oracle && python tests/golden/make_golden_config_fuzz.py        -> tests/golden/config_fuzz.json
"""
from __future__ import annotations

import json
import os
import random
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
IMAGES = HERE / "images"
sys.path.insert(0, str(HERE.parent))

import oracle_lib as O  # noqa: E402

SCALARS = ["hfov", "hang", "vang", "pos_x", "pos_y", "pos_z", "min_height", "max_height", "lum_r", "lum_g", "lum_b",
           "grid_width", "ortho_width", "step_dist", "mouse_sens", "scroll_sens", "move", "cycle", "recording_frame_count"]
TRIPLES = ["pos", "lum", "bg_color"]
NUMBERS = ["0", "1", "-1", "3", "17", "255", "256", "-0", "+5", "90", "0.5", "-2.25", ".5", "5.", "1e3", "1E-2", "-1.5e+2",
           "007", "1e999", "-1e999", "1e-999", "4294967296", "-2147483649", "99999999999999999999", "0x10", "1.5x", "12abc",
           "abc", "nan", "inf", "-inf", "1,5", "1_000", "--1", "+", "-", ".", "e5", "3.7", "2.999999999", "1e", "1e+",
           "0.1e1", "١", "1\t2"]
MAPS = ["grey8.png", "rgb8.png", "rgba8.png", "grey.jpg", "rgb24.bmp", "rgba.tga", "rle.hdr", "rgb_raw.pic", "other_size.png",
        "does_not_exist.png", "grey.pgm"]


def make(rng: random.Random) -> str:
    toks = []
    if rng.random() < 0.85:
        toks += ["heightmap", rng.choice(MAPS[:8])]
    if rng.random() < 0.85:
        toks += ["colormap", rng.choice(MAPS[:8])]
    for _ in range(rng.randint(1, 14)):
        k = rng.random()
        if k < 0.45:
            toks += [rng.choice(SCALARS), rng.choice(NUMBERS) if rng.random() < 0.8 else str(round(rng.uniform(-500, 500), rng.randint(0, 6)))]
        elif k < 0.62:
            toks += [rng.choice(TRIPLES)] + [rng.choice(NUMBERS) for _ in range(rng.choice([3, 3, 3, 2, 1, 4]))]
        elif k < 0.72:
            toks += ["resolution"] + [rng.choice(NUMBERS[:12] + ["640", "480", "3.7", "abc", "-5", "0"]) for _ in range(rng.choice([2, 2, 1, 3]))]
        elif k < 0.78:
            toks += ["print"]
        elif k < 0.86:
            toks += [rng.choice(["heightmap", "colormap"]), rng.choice(MAPS)]
        elif k < 0.93:
            toks += [rng.choice(["Hfov", "fov", "cycle_bits", "#", "// comment", "pos_w", "heightmap_", "resolution=", "=", "print;"])]
        else:
            toks += [rng.choice(SCALARS + TRIPLES + ["resolution", "heightmap", "colormap"])]      # argument missing / next id eaten
    seps = [" ", "\n", "  ", "\t", "\r\n", " \n "]
    return "".join(t + rng.choice(seps) for t in toks)


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle` first")
    rng = random.Random(20260)
    env = dict(os.environ, HMRM_FAKE_FRAMES="1", HMRM_FAKE_PROJ="1")
    out, hangs = [], 0
    with tempfile.TemporaryDirectory() as td:
        while len(out) < 400:
            text = make(rng)
            cfg = Path(td) / "c.txt"
            cfg.write_text(text)
            try:
                res = subprocess.run([str(O.REF_BIN), str(cfg)], capture_output=True, text=True, cwd=str(IMAGES), env=env, timeout=20)
            except subprocess.TimeoutExpired:
                hangs += 1          # e.g. a resolution / step the reference then renders forever: not a parser property
                continue
            if res.returncode < 0:
                hangs += 1          # the reference crashed on what it parsed (0-sized window, ...): nothing to mirror
                continue
            out.append(dict(config=text, stdout=res.stdout, stderr=res.stderr, returncode=res.returncode))
    (HERE / "config_fuzz.json").write_text(json.dumps(out, indent=0) + "\n")
    print(len(out), "cases,", hangs, "skipped (reference hung or crashed after parsing)")


if __name__ == "__main__":
    main()
