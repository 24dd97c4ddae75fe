#!/bin/bash
# Round-2 GPU session 4: tests (set-up changes, K1 dilation, packed-lane kernel), A/B timings, K1 launch list.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/s4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s4_pytest.log
tail -4 gpurun_out/s4_pytest.log
V=heightmap-ray-marcher_b200/variants
for wl in flythrough4k ortho4k spherical1080 bands8k sample720; do
  for var in base noearly; do
    echo "== $wl $var"
    if [ $var = base ]; then unset HMRM_LIBRARY; else export HMRM_LIBRARY=$PWD/$V/libhmrm_$var.so; fi
    timeout 300 python tools/profile_frame.py --workload $wl --frames 10
  done
  for var in base th8 th24; do
    echo "== $wl pack_$var"
    if [ $var = base ]; then unset HMRM_LIBRARY; else export HMRM_LIBRARY=$PWD/$V/libhmrm_$var.so; fi
    timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --traversal pack
  done
done > gpurun_out/s4_variants.txt 2>&1
unset HMRM_LIBRARY
for wl in flythrough4k ortho4k; do
  echo "== stats pack $wl"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 2 --stats --traversal pack
done > gpurun_out/s4_stats.txt 2>&1
python tools/profile_frame.py --workload flythrough4k --frames 1 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_ -c 40 --csv \
  --log-file gpurun_out/s4_k1_launches.csv python tools/profile_frame.py --workload flythrough4k --frames 1 > gpurun_out/s4_k1.log 2>&1
