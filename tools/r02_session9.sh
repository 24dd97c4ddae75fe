#!/bin/bash
# Round-2 GPU session 9 (N GPUs): band exchanges at N — fused stores RGB8 / RGBA8 against the staged copy-engine push.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_peer_frame.py -q -m gpu -x > gpurun_out/s9_pytest.log 2>&1; tail -2 gpurun_out/s9_pytest.log
for v in "peer-copy rgb8" "peer-copy rgba8" "peer rgba8" "peer rgb8"; do
  set -- $v
  timeout 600 $TR bench.py --gpus $N --steps 96 --warmup 3 --workload bands8k --exchange $1 --pixels $2 > gpurun_out/s9_bands_n${N}_$1_$2.json 2> gpurun_out/s9_bands_n${N}_$1_$2.err
  echo "bands $1 $2 exit $?"
done
