#!/bin/bash
# Round-2 GPU session 10 (N GPUs): band exchanges at N, frames in flight.
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for v in "peer-copy 2" "peer 2" "peer-copy 3"; do
  set -- $v
  timeout 600 $TR bench.py --gpus $N --steps 96 --warmup 3 --workload bands8k --exchange $1 --inflight $2 > gpurun_out/s10_bands_n${N}_$1_if$2.json 2> gpurun_out/s10_bands_n${N}_$1_if$2.err
  echo "bands $1 inflight $2 exit $?"
done
