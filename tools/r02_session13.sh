#!/bin/bash
# Round-2 GPU session 13 (N GPUs): the driver's command at N with the final build (flythrough4k + bands8k sub-record).
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/s13_bench_n$N.json 2> gpurun_out/s13_bench_n$N.err
echo "bench n$N exit $?"
tail -n 3 gpurun_out/s13_bench_n$N.err
