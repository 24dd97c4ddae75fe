"""profiles/traffic.json from ncu CSV logs: per workload, DRAM bytes (read + write) and warp instructions of ONE launch
of the production render kernel (what bench.py reports as roofline.traffic / warp_inst_per_launch).

    python tools/make_traffic_json.py gpurun_out/<prefix>_{workload}.csv ... --note "..."

Each CSV is `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,... --csv` of
`tools/profile_frame.py --workload W --frames 3 [--rgb8]`, kernel filter k2_render; the first profiled launch is skipped
when there are several (cold caches)."""
import argparse
import collections
import csv
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ap = argparse.ArgumentParser()
ap.add_argument("csvs", nargs="+")
ap.add_argument("--note", default="")
ap.add_argument("--merge", action="store_true", help="add to / replace entries of the existing profiles/traffic.json")
args = ap.parse_args()

out = {}
if args.merge:
    out = json.loads((ROOT / "profiles" / "traffic.json").read_text())
for path in args.csvs:
    m = re.search(r"_([a-z0-9]+)\.csv$", path)
    workload = m.group(1)
    lines = [l for l in open(path) if not l.startswith("==")]
    per_launch = collections.OrderedDict()
    kernel = None
    for r in csv.DictReader(lines):
        kernel = r["Kernel Name"]
        try:
            per_launch.setdefault(r["ID"], {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    launches = list(per_launch.values())
    use = launches[1:] if len(launches) > 1 else launches
    mean = lambda k: sum(l[k] for l in use) / len(use)
    out[workload] = {
        "dram_bytes": int(mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum")),
        "dram_read_bytes": int(mean("dram__bytes_read.sum")),
        "dram_write_bytes": int(mean("dram__bytes_write.sum")),
        "warp_inst": int(mean("smsp__inst_executed.sum")),
        "lanes_per_inst": round(mean("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
        "issue_active_pct": round(mean("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "kernel_us_under_ncu": round(mean("gpu__time_duration.sum") / 1e3, 1),
        "kernel": kernel,
        "launches_averaged": len(use),
        "source": path,
    }
if args.merge and "_note" in out and not args.note:
    args.note = out["_note"]
out["_note"] = args.note or "ncu --metrics (clock control none), tools/profile_frame.py, one launch of the render kernel per workload"
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
