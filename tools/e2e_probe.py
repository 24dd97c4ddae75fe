"""Streaming API (hmrm_render_async + hmrm_wait_pending) frame rate for a terrain view and for an all-sky view of the
same size: tells how much of bench.py's e2e time is the PCIe copy-out itself (all-sky: the kernel is ~0.13 ms) and how
much the kernels add."""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bench  # noqa: E402
import hmrm_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--depth", type=int, default=3)
ap.add_argument("--wc", action="store_true", help="write-combined pinned host buffers (cudaHostAllocWriteCombined)")
ap.add_argument("--rgb8", action="store_true", help="RGB8 frames (no alpha byte)")
args = ap.parse_args()
hmrm = hmrm_pkg.load()
from heightmap_ray_marcher_b200 import binding  # noqa: E402

wl = bench.WORKLOADS["flythrough4k"]
r = hmrm.Renderer(0)
r.min_height, r.max_height = bench.MIN_HEIGHT, bench.MAX_HEIGHT
r.synth_maps(wl["log2n"], bench.SEED)
ch = 3 if args.rgb8 else 4
bufs = [binding.pinned_empty((wl["H"], wl["W"], ch), write_combined=args.wc) for _ in range(args.depth + 1)]
for label, vang in (("terrain", None), ("all-sky", 30.0)):
    frames = []
    for i in range(args.frames):
        c = bench.camera(wl, i)
        frames.append(r.frame(projection=1, screen_width=wl["W"], screen_height=wl["H"], cam_pos=c["pos"],
                              hang=hmrm.deg2rad(c["hang_deg"]), vang=hmrm.deg2rad(vang if vang is not None else c["vang_deg"]),
                              hfov=hmrm.deg2rad(c["hfov_deg"]), grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"],
                              pixel_format=1 if args.rgb8 else 0))
    for rep in range(2):
        t0 = time.perf_counter()
        for i, f in enumerate(frames):
            r.render_async(f, bufs[i % len(bufs)])
            r.wait_pending(args.depth)
        r.wait()
        dt = (time.perf_counter() - t0) / len(frames)
    gbs = wl["W"] * wl["H"] * ch / dt / 1e9
    print(f"{label:8s} depth {args.depth} {'rgb8' if args.rgb8 else 'rgba8'} {'write-combined' if args.wc else 'pinned'}: "
          f"{dt * 1e3:.3f} ms/frame = {gbs:.1f} GB/s of frames")
r.close()
