"""Render a few frames of a bench workload and nothing else (short command line for ncu / compute-sanitizer)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bench  # noqa: E402
import hmrm_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="flythrough4k")
ap.add_argument("--traversal", default="skip")
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--fast", action="store_true", help="HMRM_FP32_FAST")
ap.add_argument("--vang", type=float, default=None, help="override vang (degrees), e.g. 30 = all sky")
ap.add_argument("--layout", choices=["rowmajor", "tile4", "zorder"], default=None, help="pyramid layout (default: the library's)")
ap.add_argument("--rgb8", action="store_true")
ap.add_argument("--band", default=None, help="N:I = render only rank I's interleaved tile rows of an N-way band split")
args = ap.parse_args()

hmrm = hmrm_pkg.load()
wl = bench.WORKLOADS[args.workload]
r = hmrm.Renderer(0)
r.min_height, r.max_height = bench.MIN_HEIGHT, bench.MAX_HEIGHT
if args.layout:
    r.set_layout({"rowmajor": 0, "tile4": 1, "zorder": 2}[args.layout])
r.synth_maps(wl["log2n"], bench.SEED)
band_count, band_index = (int(v) for v in args.band.split(":")) if args.band else (0, 0)
trav = {"auto": 0, "brute": 1, "skip": 2, "skip_fp64": 3, "pack": 4}[args.traversal]
for i in range(args.frames):
    c = bench.camera(wl, args.first + i)
    if args.vang is not None:
        c["vang_deg"] = args.vang
    f = r.frame(projection=wl["projection"], screen_width=wl["W"], screen_height=wl["H"], cam_pos=c["pos"],
                hang=hmrm.deg2rad(c["hang_deg"]), vang=hmrm.deg2rad(c["vang_deg"]), hfov=hmrm.deg2rad(c["hfov_deg"]),
                ortho_width=c["ortho_width"], grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"], traversal=trav, precision=1 if args.fast else 0,
                flags=hmrm.FLAG_STATS if args.stats else 0, pixel_format=1 if args.rgb8 else 0,
                band_count=band_count, band_index=band_index)
    out = r.render(f)
    st = r.stats()
    print(f"frame {args.first + i}: kernel {st.kernel_ms:.3f} ms" +
          (f" rays {st.rays} box {st.box_hits} surf {st.surf_hits} steps {st.steps} fetches {st.fetches} max {st.max_steps}"
           if args.stats else ""))
    if args.stats:
        d = r.debug_counters()
        iters = d[0] + d[2] + d[4] + d[5] + d[7]
        print("   jumps %d (covering %d samples, %.1f/jump) plain-above %d descents %d cell-above %d cell-below %d slow-locate %d"
              % (d[0], d[1], d[1] / max(d[0], 1), d[2], d[3], d[4], d[5], d[7]))
        print("   lane iterations %d, warp iterations %d: %.1f of 32 lanes busy in the march loop" % (iters, d[6], iters / max(d[6], 1)))
        if d[8] + d[9] + d[10] + d[11] == d[6] and d[6] > 0:
            print("   warp iterations by busy lanes: 1-8: %.1f %%, 9-16: %.1f %%, 17-24: %.1f %%, 25-32: %.1f %%"
                  % tuple(100.0 * d[8 + i] / d[6] for i in range(4)))
        print("   slowest tile %d clk (%.1f us at 1.965 GHz), mean tile %.0f clk, tiles > 100k clk: %d, most iterations of a ray: %d"
              % (d[8], d[8] / 1965.0, d[9] / max(st.rays / 32, 1), d[11], d[10]))
r.close()
