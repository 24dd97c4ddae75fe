#!/bin/bash
# Round-2 GPU session 16 (N GPUs): the driver's command at N with the peer protocol as one-thread kernels and as
# stream memory operations; bands8k alone with more frames in flight.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/s16_n${N}_$name.json 2> gpurun_out/s16_n${N}_$name.err; echo "$name exit $?"; }
HMRM_PEER_SYNC=kernels run bench_kernels $TR bench.py --gpus $N --steps 60 --warmup 3
HMRM_PEER_SYNC=memops  run bench_memops  $TR bench.py --gpus $N --steps 60 --warmup 3
B="bench.py --gpus $N --workload bands8k --steps 96 --warmup 3 --no-cpu-baseline"
HMRM_PEER_SYNC=memops  run bands_memops_if2 $TR $B --bands-inflight 2
HMRM_PEER_SYNC=memops  run bands_memops_if6 $TR $B --bands-inflight 6
HMRM_PEER_SYNC=kernels run bands_kernels_if6 $TR $B --bands-inflight 6
if [ "$N" -gt 2 ]; then
HMRM_PEER_SYNC=memops  run bands_memops_stores $TR $B --exchange peer
else
HMRM_PEER_SYNC=memops  run bands_memops_copy $TR $B --exchange peer-copy
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/s16_n${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    b = d.get("bands8k", d)
    pr = b.get("per_rank", {})
    print(f, "| value %.0f e2e %.0f |" % (d["value"], d["e2e"]["value"]), "bands ms/step %.4f" % b["ms_per_step"], "eff", b.get("strong_scaling_efficiency"),
          "equal", b["config"].get("gathered_frame_equals_single_gpu_frame"), "host", [round(x, 4) for x in pr.get("host_enqueue_ms_per_step", [])][:2],
          "per-rank", [round(x, 4) for x in pr.get("timed_region_ms_per_step", [])])
PY
tail -n 2 gpurun_out/s16_n${N}_*.err | grep -v "^$" | grep -i "error\|fail\|Traceback" | head
