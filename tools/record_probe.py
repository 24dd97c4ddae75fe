"""Recording throughput of the `hmap` binary (config file -> N PNG files): render + D2H + PNG encode on worker threads."""
import math
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bench  # noqa: E402
import oracle_lib as O  # noqa: E402

W, H, log2n, frames = 3840, 2160, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 48
wl = dict(bench.WORKLOADS["flythrough4k"], log2n=log2n, W=W, H=H)
hm, cm = O.synth_maps(log2n, bench.SEED)
with tempfile.TemporaryDirectory(prefix="hmrm_rec_", dir="/dev/shm") as td:
    td = Path(td)
    with open(td / "height.pgm", "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (hm.shape[1], hm.shape[0]))
        f.write(np.ascontiguousarray(hm[:, :, 0]).tobytes())
    O.write_tga_rgba(td / "color.tga", cm)
    cams = [bench.camera(wl, i) for i in range(frames)]
    kw = dict(width=W, height=H, grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"], min_height=bench.MIN_HEIGHT,
              max_height=bench.MAX_HEIGHT, **cams[0])
    (td / "config.txt").write_text(O.config_text(kw, td / "height.pgm", td / "color.tga"))
    (td / "frames.txt").write_text("".join("pos %r %r %r hang %r vang %r\n" % (*c["pos"], c["hang_deg"], c["vang_deg"]) for c in cams))
    for extra in ([], ["--frames", "1"]):
        t0 = time.perf_counter()
        subprocess.run([str(ROOT / "heightmap-ray-marcher_b200" / "hmap"), "config.txt", "--script", "frames.txt", "--out-prefix",
                        str(td / "f_")] + extra, cwd=str(td), check=True, stdout=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        n = 1 if extra else frames
        print(f"{n} frames of {W}x{H}: {dt:.2f} s wall" + ("" if extra else f" (start-up included)"))
    size = sum(p.stat().st_size for p in td.glob("f_*.png")) / frames / 1e6
    print(f"mean PNG size {size:.1f} MB")
