#!/bin/bash
# Round-2 GPU session 8 (N GPUs): the driver's line at N (flythrough4k + bands8k sub-record), band exchanges, D2H probe.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/s8_bench_n$N.json 2> gpurun_out/s8_bench_n$N.err
echo "bench n$N exit $?"
for ex in peer-allreduce gather; do
  timeout 600 $TR bench.py --gpus $N --steps 96 --warmup 3 --workload bands8k --exchange $ex > gpurun_out/s8_bands_n${N}_$ex.json 2> gpurun_out/s8_bands_n${N}_$ex.err
  echo "bands $ex exit $?"
done
timeout 300 $TR tools/d2h_probe.py > gpurun_out/s8_d2h_probe_n$N.txt 2>&1
tail -n 3 gpurun_out/s8_bench_n$N.err
