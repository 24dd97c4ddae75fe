#!/bin/bash
# Round-2 GPU session 3 (2 GPUs): the driver's N=2 line (flythrough4k + bands8k sub-record), band exchanges side by side.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/s3_bench_n$N.json 2> gpurun_out/s3_bench_n$N.err
echo "bench n$N exit $?"
for ex in peer peer-allreduce gather; do
  timeout 600 $TR bench.py --gpus $N --steps 60 --warmup 3 --workload bands8k --exchange $ex > gpurun_out/s3_bands_n${N}_$ex.json 2> gpurun_out/s3_bands_n${N}_$ex.err
  echo "bands $ex exit $?"
done
timeout 600 $TR bench.py --gpus $N --steps 60 --warmup 3 --workload bands8k --exchange peer --inflight 1 > gpurun_out/s3_bands_n${N}_peer_if1.json 2> gpurun_out/s3_bands_n${N}_peer_if1.err
timeout 600 $TR bench.py --gpus $N --steps 60 --warmup 3 --workload bands8k --exchange peer --pixels rgba8 > gpurun_out/s3_bands_n${N}_peer_rgba.json 2> gpurun_out/s3_bands_n${N}_peer_rgba.err
timeout 600 python bench.py --steps 120 --warmup 3 --no-cpu-baseline --pixels rgba8 > gpurun_out/s3_bench_n1_rgba.json 2> gpurun_out/s3_bench_n1_rgba.err
timeout 600 python bench.py --steps 120 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench_n1_rgb.json 2> gpurun_out/s3_bench_n1_rgb.err
tail -3 gpurun_out/s3_*.err
