#!/bin/bash
# Round-2 GPU session 14 (8 GPUs): where the 8K band split loses time at N = 8 — slot reset by kernel vs memsets,
# no exchange at all (the render kernels' own floor), frames in flight.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
B="bench.py --gpus $N --workload bands8k --steps 96 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/s14_$name.json 2> gpurun_out/s14_$name.err; echo "$name exit $?"; }
run default            $TR $B
HMRM_RESET_MEMSET=1 run memset $TR $B
run none_if3           $TR $B --exchange none
run none_if1           $TR $B --exchange none --bands-inflight 1
run if4                $TR $B --bands-inflight 4
run if2                $TR $B --bands-inflight 2
run stores             $TR $B --exchange peer
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s14_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    pr = d.get("per_rank", {})
    print(f, "ms/step %.4f" % d["ms_per_step"], "eff", d.get("strong_scaling_efficiency"), "host", d.get("host_enqueue_ms_per_step"),
          "per-rank", [round(x, 4) for x in pr.get("timed_region_ms_per_step", [])], "k alone", [round(x, 4) for x in pr.get("render_kernel_alone_ms", [])])
PY
