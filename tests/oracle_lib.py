"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/_build/liboracle.so)
and runners for the unmodified reference build (oracle/_ref/hmap_ref, ref_harness).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  Nothing here is on the product path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ORACLE_DIR = REPO / "oracle"
ORACLE_SO = ORACLE_DIR / "_build" / "liboracle.so"
REF_DIR = ORACLE_DIR / "_ref"
REF_BIN = REF_DIR / "hmap_ref"
REF_BIN_O2 = REF_DIR / "hmap_ref_O2"
REF_HARNESS = REF_DIR / "ref_harness"

PERSPECTIVE, SPHERICAL, ORTHOGRAPHIC = 1, 2, 3


class OracleFrame(C.Structure):
    """Mirror of `oracle_frame` (oracle/hmap_oracle.h) = the reference's render globals."""

    _fields_ = [
        ("projection", C.c_int32),
        ("screen_width", C.c_int32),
        ("screen_height", C.c_int32),
        ("cam_pos", C.c_double * 3),
        ("hang", C.c_double),
        ("vang", C.c_double),
        ("hfov", C.c_double),
        ("ortho_width", C.c_double),
        ("grid_width", C.c_double),
        ("step_dist", C.c_double),
        ("min_height", C.c_double),
        ("max_height", C.c_double),
        ("bg", C.c_uint8 * 3),
        ("pad_", C.c_uint8),
        ("cycle", C.c_int32),
        ("cycle_period", C.c_int32),
    ]


class OracleStats(C.Structure):
    _fields_ = [
        ("rays", C.c_int64),
        ("box_hits", C.c_int64),
        ("surf_hits", C.c_int64),
        ("steps", C.c_int64),
        ("max_steps", C.c_int64),
        ("status", C.c_int32),
    ]


_lib = None


def build_oracle() -> None:
    subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()          # make: a no-op when oracle/_build/liboracle.so is newer than its sources
        L = C.CDLL(str(ORACLE_SO))
        L.oracle_deg2rad.restype = C.c_double
        L.oracle_deg2rad.argtypes = [C.c_double]
        L.oracle_update_heightmap.restype = None
        L.oracle_update_heightmap.argtypes = [C.c_void_p, C.c_int64] + [C.c_double] * 5 + [C.c_void_p]
        L.oracle_camera_basis.restype = None
        L.oracle_camera_basis.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_get_ray.restype = None
        L.oracle_get_ray.argtypes = [C.POINTER(OracleFrame), C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_distance.restype = C.c_double
        L.oracle_distance.argtypes = [C.c_void_p] * 4
        L.oracle_intersection.restype = C.c_int
        L.oracle_intersection.argtypes = [C.c_void_p] * 5
        L.oracle_render.restype = C.c_int
        L.oracle_render.argtypes = [C.POINTER(OracleFrame), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(OracleStats)]
        L.oracle_render_synth.restype = C.c_int
        L.oracle_render_synth.argtypes = [C.POINTER(OracleFrame), C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                          C.POINTER(OracleStats)]
        L.oracle_synth_maps.restype = None
        L.oracle_synth_maps.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def deg2rad(deg: float) -> float:
    return lib().oracle_deg2rad(float(deg))


def make_frame(projection=PERSPECTIVE, width=800, height=600, pos=(-5.0, 5.0, 0.0), hang_deg=None,
               vang_deg=None, hfov_deg=None, hang=None, vang=None, hfov=None, ortho_width=0.1,
               grid_width=0.05, step_dist=0.25, min_height=0.0, max_height=10.0, bg=(0, 0, 0),
               cycle=0, cycle_period=1) -> OracleFrame:
    """Defaults are the reference's (main/hmap.cpp:31-112) except cycle_period (1 = full frame).

    Angles given in degrees go through the reference's DegreesToRads rounding
    (main/hmap.cpp:131-133) exactly as its config parser does."""
    f = OracleFrame()
    f.projection = projection
    f.screen_width, f.screen_height = width, height
    f.cam_pos[:] = [float(v) for v in pos]
    f.hang = deg2rad(hang_deg) if hang_deg is not None else (hang if hang is not None else -math.pi / 4.0)
    f.vang = deg2rad(vang_deg) if vang_deg is not None else (vang if vang is not None else math.pi / 2.0)
    f.hfov = deg2rad(hfov_deg) if hfov_deg is not None else (hfov if hfov is not None else math.pi / 2.0)
    f.ortho_width = ortho_width
    f.grid_width = grid_width
    f.step_dist = step_dist
    f.min_height = min_height
    f.max_height = max_height
    f.bg[:] = [int(v) & 255 for v in bg]
    f.cycle = cycle
    f.cycle_period = cycle_period
    return f


def update_heightmap(rgb8: np.ndarray, lum=(0.299, 0.587, 0.114), min_height=0.0, max_height=10.0) -> np.ndarray:
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w = rgb8.shape[:2]
    out = np.empty((h, w), dtype=np.float64)
    lib().oracle_update_heightmap(rgb8.ctypes.data, h * w, lum[0], lum[1], lum[2], min_height, max_height,
                                  out.ctypes.data)
    return out


def render(frame: OracleFrame, heights: np.ndarray, colormap: np.ndarray, want_steps=True,
           rows=None, framebuf=None):
    """Returns (framebuf RGBA8 [H,W,4], step_index int32 [H,W] or None, OracleStats)."""
    heights = np.ascontiguousarray(heights, dtype=np.float64)
    colormap = np.ascontiguousarray(colormap, dtype=np.uint8)
    mh, mw = heights.shape
    assert colormap.shape == (mh, mw, 4)
    W, H = frame.screen_width, frame.screen_height
    if framebuf is None:
        framebuf = np.zeros((H, W, 4), dtype=np.uint8)
    steps = np.full((H, W), -3, dtype=np.int32) if want_steps else None
    st = OracleStats()
    r0, r1 = rows if rows is not None else (0, H)
    lib().oracle_render(C.byref(frame), heights.ctypes.data, colormap.ctypes.data, mw, mh,
                        framebuf.ctypes.data, steps.ctypes.data if want_steps else None, r0, r1, C.byref(st))
    return framebuf, steps, st


def render_synth(frame: OracleFrame, log2n: int, seed: int = 1234, lum=(0.299, 0.587, 0.114), rows=None,
                 want_steps=True):
    """oracle_render over the procedural synthetic map of size 2^log2n, texels generated on the fly (no arrays):
    full-size parity bands for the BASELINE map sizes.  Returns (framebuf, step_index, stats); only `rows` are written."""
    W, H = frame.screen_width, frame.screen_height
    framebuf = np.zeros((H, W, 4), dtype=np.uint8)
    steps = np.full((H, W), -3, dtype=np.int32) if want_steps else None
    st = OracleStats()
    r0, r1 = rows if rows is not None else (0, H)
    lib().oracle_render_synth(C.byref(frame), log2n, seed, lum[0], lum[1], lum[2], framebuf.ctypes.data,
                              steps.ctypes.data if want_steps else None, r0, r1, C.byref(st))
    return framebuf, steps, st


def get_ray(frame: OracleFrame, w: float, h: float):
    pos = (C.c_double * 3)()
    d = (C.c_double * 3)()
    lib().oracle_get_ray(C.byref(frame), w, h, pos, d)
    return tuple(pos), tuple(d)


def distance(pos, d, c0, c1):
    a = [(C.c_double * 3)(*v) for v in (pos, d, c0, c1)]
    return lib().oracle_distance(*a)


def intersection(pos, d, c0, c1):
    a = [(C.c_double * 3)(*v) for v in (pos, d, c0, c1)]
    out = (C.c_double * 3)(0.0, 0.0, 0.0)
    hit = lib().oracle_intersection(out, *a)
    return bool(hit), tuple(out)


def synth_maps(log2n: int, seed: int = 1234):
    """Integer fBm height (RGB8, R=G=B) + RGBA8 colormap of size 2^log2n (csrc/synth_fbm.h)."""
    n = 1 << log2n
    hm = np.empty((n, n, 3), dtype=np.uint8)
    cm = np.empty((n, n, 4), dtype=np.uint8)
    lib().oracle_synth_maps(log2n, seed, hm.ctypes.data, cm.ctypes.data)
    return hm, cm


# --------------------------------------------------------------------------------------
# image files the reference's stb_image can read (vendor/stb_image.h: PNG, PNM, TGA ...)
# --------------------------------------------------------------------------------------
def write_png(path, arr: np.ndarray) -> None:
    from PIL import Image

    mode = {3: "RGB", 4: "RGBA"}[arr.shape[2]] if arr.ndim == 3 else "L"
    Image.fromarray(arr, mode).save(path, compress_level=1)


def write_ppm(path, rgb8: np.ndarray) -> None:
    h, w = rgb8.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb8[:, :, :3]).tobytes())


def write_tga_rgba(path, rgba8: np.ndarray) -> None:
    """Uncompressed 32-bit true-colour TGA, top-left origin (BGRA byte order on disk)."""
    h, w = rgba8.shape[:2]
    hdr = bytearray(18)
    hdr[2] = 2
    hdr[12:14] = int(w).to_bytes(2, "little")
    hdr[14:16] = int(h).to_bytes(2, "little")
    hdr[16] = 32
    hdr[17] = 0x28
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.ascontiguousarray(rgba8[:, :, [2, 1, 0, 3]]).tobytes())


# --------------------------------------------------------------------------------------
# the unmodified reference (Oracle A) under the fake SDL backend
# --------------------------------------------------------------------------------------
def have_ref() -> bool:
    return REF_BIN.exists() and REF_HARNESS.exists()


def config_text(frame_kw: dict, heightmap_path, colormap_path, lum=None, extra="") -> str:
    """Config file in the reference's grammar (main/hmap.cpp:309-489).  Angles in degrees."""
    k = frame_kw
    lines = [
        f"resolution {k['width']} {k['height']}",
        f"hfov {k.get('hfov_deg', 90)!r}",
        f"hang {k.get('hang_deg', -45)!r}",
        f"vang {k.get('vang_deg', 90)!r}",
        "pos " + " ".join(repr(float(v)) for v in k.get("pos", (-5.0, 5.0, 0.0))),
        f"min_height {k.get('min_height', 0.0)!r}",
        f"max_height {k.get('max_height', 10.0)!r}",
        f"grid_width {k.get('grid_width', 0.05)!r}",
        f"ortho_width {k.get('ortho_width', 0.1)!r}",
        f"step_dist {k.get('step_dist', 0.25)!r}",
        "bg_color " + " ".join(str(int(v)) for v in k.get("bg", (0, 0, 0))),
        "cycle 1",
    ]
    if lum is not None:
        lines.append("lum " + " ".join(repr(float(v)) for v in lum))
    lines.append(f"heightmap {heightmap_path}")
    lines.append(f"colormap {colormap_path}")
    if extra:
        lines.append(extra)
    return "\n".join(lines) + "\n"


def run_ref(config: str, projection: int, width: int, height: int, frames=1, script=None, binary=None,
            threads=None, timeout=3600, warmup=0):
    """Run the unmodified reference headlessly; returns (list of RGBA8 frames, list of ms)."""
    binary = Path(binary) if binary else REF_BIN
    with tempfile.TemporaryDirectory(prefix="hmrm_ref_") as td:
        cfg = Path(td) / "config.txt"
        cfg.write_text(config)
        env = dict(os.environ)
        env["HMRM_FAKE_PROJ"] = str(projection)
        env["HMRM_FAKE_FRAMES"] = str(frames)
        env["HMRM_FAKE_WARMUP"] = str(warmup)
        env["HMRM_FAKE_DUMP"] = str(Path(td) / "frame_")
        env["HMRM_FAKE_TIMES"] = str(Path(td) / "times.txt")
        if script is not None:
            sp = Path(td) / "script.txt"
            sp.write_text("\n".join(script) + "\n")
            env["HMRM_FAKE_SCRIPT"] = str(sp)
            frames = len(script)
        if threads is not None:
            env["OMP_NUM_THREADS"] = str(threads)
        subprocess.run([str(binary), str(cfg)], check=True, env=env, cwd=td, timeout=timeout,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out = []
        for i in range(frames):
            raw = np.fromfile(Path(td) / f"frame_{i}.rgba", dtype=np.uint8)
            out.append(raw.reshape(height, width, 4))
        times = [float(l.split()[1]) for l in (Path(td) / "times.txt").read_text().split("\n") if l.strip()]
        return out, times


def ref_heights(rgb8: np.ndarray, lum, min_height, max_height) -> np.ndarray:
    """UpdateHeightmap of the unmodified reference (through oracle/_ref/ref_harness)."""
    h, w = rgb8.shape[:2]
    with tempfile.TemporaryDirectory(prefix="hmrm_ref_") as td:
        src = Path(td) / "in.raw"
        dst = Path(td) / "out.f64"
        np.ascontiguousarray(rgb8, dtype=np.uint8).tofile(src)
        subprocess.run([str(REF_HARNESS), "heights", str(src), str(w), str(h)] +
                       [repr(float(v)) for v in (*lum, min_height, max_height)] + [str(dst)], check=True)
        return np.fromfile(dst, dtype=np.float64).reshape(h, w)


def ref_kat(queries: list[str]) -> list[str]:
    """Feed hex-float queries to ref_harness kat; returns one answer line per query."""
    res = subprocess.run([str(REF_HARNESS), "kat"], input="\n".join(queries) + "\n", text=True,
                         capture_output=True, check=True)
    return [l for l in res.stdout.split("\n") if l.strip()]
