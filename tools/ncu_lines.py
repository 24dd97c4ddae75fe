"""Attribute an ncu report's per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k2_render_skipILb0 [--top 40]

ncu's CSV export of the source page carries metrics only in the SASS view; this joins it with
`nvdisasm -g` line information of the same cubin (extracted from libhmrm.so, which must be the build
that was profiled) by instruction order inside the kernel.
"""
from __future__ import annotations

import argparse
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "heightmap-ray-marcher_b200" / "libhmrm.so"


def sass_lines(kernel_substr: str):
    """-> list of (file, line) per SASS instruction of the first function whose name contains kernel_substr."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubins = list(Path(td).glob("*.cubin"))
        text = ""
        for cb in cubins:
            text += subprocess.run(["nvdisasm", "-g", "-c", str(cb)], capture_output=True, text=True).stdout
    out, cur, active = [], ("?", 0), False
    for ln in text.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if active and out:
                break
            active = kernel_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()

    csv_text = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True,
                              text=True).stdout
    rows = list(csv.reader(io.StringIO(csv_text)))
    # several kernels may be in the report; take the first block whose name matches
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    base = args.kernel.split("ILb")[0]
    blk = next(b for b in blocks if base in b["name"])
    hdr = blk["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    lines = sass_lines(args.kernel)
    if len(lines) != len(blk["rows"]):
        print(f"warning: {len(lines)} SASS instructions in the cubin vs {len(blk['rows'])} in the report", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0, 0])
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    stall_by_line = defaultdict(lambda: defaultdict(int))
    for i, r in enumerate(blk["rows"]):
        key = lines[i] if i < len(lines) else ("?", 0)
        a = agg[key]
        a[0] += int(r[col["Instructions Executed"]] or 0)
        a[1] += int(r[col["Thread Instructions Executed"]] or 0)
        a[2] += int(r[col["# Samples"]] or 0)
        a[3] += 1
        for h in stall_cols:
            v = int(r[col[h]] or 0)
            if v:
                stall_by_line[key][h] += v
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[2] for a in agg.values()) or 1
    src_cache = {}

    def src(file, line):
        if file not in src_cache:
            cands = list((ROOT / "heightmap-ray-marcher_b200" / "csrc").glob(file))
            src_cache[file] = cands[0].read_text().split("\n") if cands else []
        t = src_cache[file]
        return t[line - 1].strip()[:90] if 0 < line <= len(t) else ""

    print(f"kernel {blk['name']}: {tot_i} warp instructions, {tot_s} samples, {len(blk['rows'])} SASS instructions")
    print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'thr/inst':>8s} {'sass':>5s}  top stall  | source")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: args.top]:
        st = stall_by_line[key]
        top = max(st.items(), key=lambda kv: kv[1])[0].replace("stall_", "") if st else ""
        print(f"{key[0][:20]}:{key[1]:<6d} {a[0] / tot_i * 100:6.2f} {a[2] / tot_s * 100:6.2f} "
              f"{(a[1] / a[0]) if a[0] else 0:8.1f} {a[3]:5d}  {top:10s} | {src(*key)}")


if __name__ == "__main__":
    main()
