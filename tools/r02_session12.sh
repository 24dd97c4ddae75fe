#!/bin/bash
# Round-2 GPU session 12 (final, one GPU): test suite, the driver's two arms, launch list, the other workloads' lines.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s12_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s12_pytest.log
tail -3 gpurun_out/s12_pytest.log
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/s12_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/s12_smoke.log
timeout 900 python bench.py > gpurun_out/s12_bench_ours.json 2> gpurun_out/s12_bench_ours.err; echo "ours exit $?"
timeout 900 python bench.py --impl reference > gpurun_out/s12_bench_reference.json 2> gpurun_out/s12_bench_reference.err; echo "reference exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s12_bench_ours_20steps.json 2> gpurun_out/s12_bench_ours_20steps.err
for wl in sample720 spherical1080 ortho4k bands8k; do
  timeout 900 python bench.py --workload $wl --steps 60 --warmup 3 > gpurun_out/s12_bench_$wl.json 2> gpurun_out/s12_bench_$wl.err
  echo "bench $wl exit $?"
done
python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/s12_plain.json 2> gpurun_out/s12_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s12_bench_launches.csv \
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/s12_ncu_bench.log 2>&1
