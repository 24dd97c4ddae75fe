#!/bin/bash
# Round-2 GPU session 2: full GPU test suite again, then kernel variants (HMRM_LIBRARY) on four workloads.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s2_pytest.log
tail -5 gpurun_out/s2_pytest.log
V=heightmap-ray-marcher_b200/variants
for wl in flythrough4k ortho4k spherical1080 bands8k; do
  for var in base pf1 pf2 pf3 pf2l2 t128 c3; do
    echo "== $wl $var"
    if [ $var = base ]; then unset HMRM_LIBRARY; else export HMRM_LIBRARY=$PWD/$V/libhmrm_$var.so; fi
    timeout 300 python tools/profile_frame.py --workload $wl --frames 10
  done
  unset HMRM_LIBRARY
  echo "== $wl base_g32"
  HMRM_L2_FETCH_GRANULARITY=32 timeout 300 python tools/profile_frame.py --workload $wl --frames 10
  echo "== $wl base_g128"
  HMRM_L2_FETCH_GRANULARITY=128 timeout 300 python tools/profile_frame.py --workload $wl --frames 10
done > gpurun_out/s2_variants.txt 2>&1
unset HMRM_LIBRARY
for wl in flythrough4k ortho4k; do
  echo "== stats $wl"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 3 --stats
done > gpurun_out/s2_stats.txt 2>&1
ls -la gpurun_out | tail -5
