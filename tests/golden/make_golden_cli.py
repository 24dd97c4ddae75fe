"""Golden fixtures for the host side (config grammar + image ingest), produced by the UNMODIFIED reference.

    make -C oracle && python tests/golden/make_golden_cli.py

Outputs (committed):
    tests/golden/images/*            small test images in every PNG flavour + PGM/PPM/TGA (written with PIL / by hand)
    tests/golden/images.json         per image: what stb_image (vendor/stb_image.h, through oracle/_ref/ref_harness
                                     decode = stbi_load(path,&w,&h,&n,comp)) returns for comp 3 and 4: size + sha256
    tests/golden/config_echo.json    per config text: stdout / stderr / exit status of oracle/_ref/hmap_ref
                                     (the reference's ConsumeConfigStream echo, main/hmap.cpp:309-520)
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent
IMAGES = HERE / "images"
sys.path.insert(0, str(HERE.parent))

import oracle_lib as O  # noqa: E402


def make_images():
    IMAGES.mkdir(exist_ok=True)
    rng = np.random.RandomState(7)
    w, h = 19, 13
    rgb = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    rgba = np.dstack([rgb, rng.randint(0, 256, size=(h, w)).astype(np.uint8)])
    rgba[::3, ::2, 3] = 0
    grey = rng.randint(0, 256, size=(h, w)).astype(np.uint8)
    Image.fromarray(rgb, "RGB").save(IMAGES / "rgb8.png")
    Image.fromarray(rgba, "RGBA").save(IMAGES / "rgba8.png")
    Image.fromarray(grey, "L").save(IMAGES / "grey8.png")
    Image.fromarray(np.dstack([grey, rgba[..., 3]]), "LA").save(IMAGES / "grey_alpha8.png")
    Image.fromarray(rgb, "RGB").save(IMAGES / "rgb8_interlaced.png", interlace=True) if False else None
    pal = Image.fromarray(rgb, "RGB").quantize(colors=17)
    pal.save(IMAGES / "palette.png")
    pal.save(IMAGES / "palette_trns.png", transparency=3)
    Image.fromarray((grey.astype(np.uint16) * 257 + 13).astype(np.uint16)).save(IMAGES / "grey16.png")
    Image.fromarray(grey, "L").save(IMAGES / "grey8_trns.png", transparency=int(grey[2, 3]))
    Image.fromarray(rgb, "RGB").save(IMAGES / "rgb8_trns.png", transparency=tuple(int(v) for v in rgb[1, 1]))
    for bits in (1, 2, 4):
        g = (grey >> (8 - bits)).astype(np.uint8)
        im = Image.fromarray(g * (255 // ((1 << bits) - 1)), "L")
        im.save(IMAGES / f"grey{bits}.png", bits=bits)
    # 16-bit RGB and Adam7 are not writable with PIL's PNG plugin: written by hand below
    write_png_raw(IMAGES / "rgb16.png", (rgb.astype(np.uint16) * 251 + 77).astype(np.uint16), 16, 2)
    write_png_raw(IMAGES / "rgba8_adam7.png", rgba, 8, 6, interlace=True)
    O.write_ppm(IMAGES / "rgb.ppm", rgb)
    with open(IMAGES / "grey.pgm", "wb") as f:
        f.write(b"P5\n# a comment\n%d %d\n255\n" % (w, h))
        f.write(grey.tobytes())
    O.write_tga_rgba(IMAGES / "rgba.tga", rgba)
    # every other format stb_image reads for the reference: JPEG, BMP, GIF, TGA flavours, PSD, 16-bit PNM
    subprocess.run([sys.executable, str(HERE / "gen_format_images.py"), str(IMAGES)], check=True)


def write_png_raw(path, arr, depth, color, interlace=False):
    """Minimal PNG writer for the flavours PIL cannot emit (16-bit RGB, Adam7)."""
    import struct
    import zlib

    h, w = arr.shape[:2]
    ch = arr.shape[2] if arr.ndim == 3 else 1

    def rows(sub):
        out = b""
        for r in sub:
            data = r.astype(">u2").tobytes() if depth == 16 else r.astype(np.uint8).tobytes()
            out += b"\x00" + data
        return out

    if interlace:
        xo, yo, xs, ys = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1), (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
        raw = b""
        for p in range(7):
            sub = arr[yo[p]::ys[p], xo[p]::xs[p]]
            if sub.shape[0] and sub.shape[1]:
                raw += rows(sub.reshape(sub.shape[0], -1, ch).reshape(sub.shape[0], -1))
    else:
        raw = rows(arr.reshape(h, -1))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 1 if interlace else 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def golden_images():
    meta = {}
    with tempfile.TemporaryDirectory() as td:
        for p in sorted(IMAGES.iterdir()):
            rec = {}
            for comp in (3, 4):
                out = Path(td) / "o.raw"
                res = subprocess.run([str(O.REF_HARNESS), "decode", str(p), str(comp), str(out)], capture_output=True,
                                     text=True, check=True).stdout.split()
                if res[0] == "FAIL":
                    rec[str(comp)] = None
                else:
                    rec[str(comp)] = dict(width=int(res[0]), height=int(res[1]),
                                          sha256=hashlib.sha256(out.read_bytes()).hexdigest())
            meta[p.name] = rec
            print(p.name, rec["3"] and rec["3"]["sha256"][:12], rec["4"] and rec["4"]["sha256"][:12])
    (HERE / "images.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")


CONFIGS = {
    # sample_config.txt of the reference with runnable paths: `cycle_bits 6` is not an identifier (two warnings)
    "sample": "resolution 800 450\nhfov 90\nmin_height 0.0\nmax_height 10.0\ngrid_width 0.01\northo_width 0.1\n"
              "step_dist 0.05\nbg_color 0 0 0\ncycle_bits 6\nmouse_sens 0.00003\nheightmap grey8.png\ncolormap rgba8.png\n",
    "everything": "heightmap rgb8.png colormap rgba8.png resolution 64 36 hfov 75.5 hang -30 vang 101.25 pos 1 -2 3.5 "
                  "pos_x 7 pos_y -8.25 pos_z 9e-3 min_height -1 max_height 2.5 lum 0.2 0.7 0.1 lum_norm 1 2 5 lum_r 0.3 "
                  "lum_g 0.59 lum_b 0.11 grid_width 0.02 ortho_width 0.04 step_dist 1e-2 bg_color 300 -1 17 cycle 1 "
                  "mouse_sens 2 scroll_sens 3 move 0.5 recording_frame_count 12 print\n",
    "last_wins": "hfov 10 hfov 20 heightmap grey8.png heightmap rgb8.png colormap rgba8.png hang 400.125\n",
    "bad_number": "hfov abc vang 100 heightmap grey8.png colormap rgba8.png\n",
    "missing_heightmap": "colormap rgba8.png hfov 60\n",
    "missing_colormap": "heightmap grey8.png\n",
    "unreadable_heightmap": "hfov 60 heightmap does_not_exist.png colormap rgba8.png\n",
    "size_conflict": "heightmap grey8.png colormap other_size.png\n",
    "pnm_tga": "heightmap grey.pgm colormap rgba.tga resolution 32 18 cycle 1\n",
    # the reference's own sample_config.txt names JPEG maps ("heightmap path/to/img.jpg")
    "jpeg_maps": "heightmap grey.jpg colormap base_444.jpg resolution 32 18\n",
    "bmp_gif_maps": "heightmap rgb24.bmp colormap transparent.gif\n",
    "hdr_pic_maps": "heightmap rle.hdr colormap rgb_plus_alpha_raw.pic\n",
}


def golden_configs():
    Image.fromarray(np.zeros((5, 7, 4), dtype=np.uint8), "RGBA").save(IMAGES / "other_size.png")
    out = {}
    env = dict(os.environ, HMRM_FAKE_FRAMES="1", HMRM_FAKE_PROJ="1")
    with tempfile.TemporaryDirectory() as td:
        for name, text in CONFIGS.items():
            cfg = Path(td) / f"{name}.txt"
            cfg.write_text(text)
            res = subprocess.run([str(O.REF_BIN), str(cfg)], capture_output=True, text=True, cwd=str(IMAGES), env=env,
                                 timeout=120)
            out[name] = dict(config=text, stdout=res.stdout, stderr=res.stderr, returncode=res.returncode)
            print(f"{name:22s} rc={res.returncode} stdout {len(res.stdout.splitlines())} lines, stderr {res.stderr.strip()[:80]!r}")
    (HERE / "config_echo.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle` first")
    make_images()
    golden_images()
    golden_configs()
