#!/bin/bash
# Round-2 GPU session 5: packed-lane kernel v3 — tests, timings against k2_render_lin, lane statistics.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/s5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s5_pytest.log
tail -4 gpurun_out/s5_pytest.log
V=heightmap-ray-marcher_b200/variants
for wl in flythrough4k ortho4k spherical1080 bands8k sample720; do
  echo "== $wl skip"
  unset HMRM_LIBRARY
  timeout 300 python tools/profile_frame.py --workload $wl --frames 10
  for var in base th8 th24 pc3; do
    echo "== $wl pack_$var"
    if [ $var = base ]; then unset HMRM_LIBRARY; else export HMRM_LIBRARY=$PWD/$V/libhmrm_$var.so; fi
    timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --traversal pack
  done
  unset HMRM_LIBRARY
  echo "== $wl pack_rgb8"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --traversal pack --rgb8
  echo "== $wl skip_rgb8"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --rgb8
done > gpurun_out/s5_variants.txt 2>&1
unset HMRM_LIBRARY
for wl in flythrough4k ortho4k; do
  echo "== stats pack $wl"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 2 --stats --traversal pack
done > gpurun_out/s5_stats.txt 2>&1
