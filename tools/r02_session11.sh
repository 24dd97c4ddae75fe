#!/bin/bash
# Round-2 GPU session 11: sky-tile batch size of the tile queue (HMRM_SKY_BATCH), single frames and the pipelined bench.
mkdir -p gpurun_out
for b in 8 16 32 64; do
  for wl in flythrough4k spherical1080 bands8k; do
    echo "== $wl batch$b"
    HMRM_SKY_BATCH=$b timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --rgb8
  done
  echo "== allsky batch$b"
  HMRM_SKY_BATCH=$b timeout 300 python tools/profile_frame.py --workload flythrough4k --frames 10 --rgb8 --vang 30
done > gpurun_out/s11_sky_batch.txt 2>&1
for b in 8 32; do
  HMRM_SKY_BATCH=$b timeout 600 python bench.py --steps 120 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench_batch$b.json 2> gpurun_out/s11_bench_batch$b.err
done
