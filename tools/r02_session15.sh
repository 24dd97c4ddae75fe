#!/bin/bash
# Round-2 GPU session 15 (1 GPU): full GPU suite on the build with the slot-reset kernel, eight launch slots and the
# stream-memory-operation peer protocol; flythrough4k with the reset kernel vs the two memsets.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s15_pytest.txt 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/s15_pytest.txt
timeout 300 python bench.py --steps 240 --warmup 3 --no-cpu-baseline > gpurun_out/s15_bench_reset_kernel.json 2> gpurun_out/s15_bench_reset_kernel.err; echo "bench exit $?"
HMRM_RESET_MEMSET=1 timeout 300 python bench.py --steps 240 --warmup 3 --no-cpu-baseline > gpurun_out/s15_bench_reset_memset.json 2> gpurun_out/s15_bench_reset_memset.err; echo "bench exit $?"
python - <<'PY'
import json
for n in ("reset_kernel", "reset_memset"):
    d = json.loads(open(f"gpurun_out/s15_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "host enqueue", d.get("host_enqueue_ms_per_step"))
PY
