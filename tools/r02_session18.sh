#!/bin/bash
# Round-2 GPU session 18 (1 GPU): the final build — GPU suite, smoke(), the driver's bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s18_pytest.txt 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/s18_pytest.txt
timeout 300 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/s18_bench_ours.json 2> gpurun_out/s18_bench_ours.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s18_bench_ours.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("parity_of_sample"), "clocks", d["clocks"])
PY
