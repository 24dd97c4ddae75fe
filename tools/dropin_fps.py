"""Frame time of the reference's OWN program (SDL loop, console, recording, everything) with and without the
INTEGRATION.md patch: oracle/_ref/hmap_ref_O2 (unmodified, -O2, all host cores) against oracle/_ref/hmap_patched (the
same program, per-pixel loop and height prepass replaced by libhmrm.so calls).  Both run under the fake SDL, which
times every frame between SDL_GetModState (main/hmap.cpp:929) and SDL_UpdateTexture (:1082); the frames are compared."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bench  # noqa: E402
import oracle_lib as O  # noqa: E402

import tempfile  # noqa: E402

for name, (log2n, W, H, frames) in {"1280x720 over 1024^2 (BASELINE configs[0])": (10, 1280, 720, 12),
                                    "3840x2160 over 4096^2": (12, 3840, 2160, 4)}.items():
    wl = dict(bench.WORKLOADS["flythrough4k"], log2n=log2n, W=W, H=H)
    hm, cm = O.synth_maps(log2n, bench.SEED)
    with tempfile.TemporaryDirectory(prefix="hmrm_dropin_") as td:
        hp, cp = Path(td) / "height.pgm", Path(td) / "color.tga"
        with open(hp, "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (hm.shape[1], hm.shape[0]))
            f.write(np.ascontiguousarray(hm[:, :, 0]).tobytes())
        O.write_tga_rgba(cp, cm)
        cams = [bench.camera(wl, 20 * i) for i in range(frames + 1)]
        kw = dict(width=W, height=H, grid_width=bench.GRID_WIDTH, step_dist=wl["step_dist"], min_height=bench.MIN_HEIGHT,
                  max_height=bench.MAX_HEIGHT, **cams[0])
        cfg = O.config_text(kw, hp, cp)
        script = ["pos %r %r %r hang %r vang %r" % (*c["pos"], c["hang_deg"], c["vang_deg"]) for c in cams]
        out = {}
        for label, binary in (("reference -O2", O.REF_BIN_O2), ("patched (libhmrm.so)", ROOT / "oracle" / "_ref" / "hmap_patched")):
            fr, ms = O.run_ref(cfg, 1, W, H, script=script, binary=binary)
            out[label] = (fr[1:], ms[1:])
        same = all(np.array_equal(a, b) for a, b in zip(out["reference -O2"][0], out["patched (libhmrm.so)"][0]))
        a, b = np.mean(out["reference -O2"][1]), np.mean(out["patched (libhmrm.so)"][1])
        print(f"{name}: reference {a:.1f} ms/frame, patched {b:.2f} ms/frame ({a / b:.0f}x), frames identical: {same}")
