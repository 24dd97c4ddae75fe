#!/bin/bash
# Round-2 GPU session 1: GPU test suite, pyramid-layout A/B (timings + ncu counters), K1 launch list, short bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s1_gpus.txt 2>&1
free -g >> gpurun_out/s1_gpus.txt 2>&1
nproc >> gpurun_out/s1_gpus.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log

for wl in flythrough4k ortho4k spherical1080 sample720 bands8k; do
  for L in rowmajor tile4 zorder; do
    echo "== $wl $L"
    timeout 300 python tools/profile_frame.py --workload $wl --layout $L --frames 8
  done
done > gpurun_out/s1_layout_times.txt 2>&1

timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench exit $?"

M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active
for wl in flythrough4k ortho4k bands8k; do
  for L in rowmajor tile4 zorder; do
    timeout 600 ncu --metrics $M --clock-control none -k regex:k2_render_lin -s 1 -c 2 --csv \
      --log-file gpurun_out/s1_ncu_${wl}_${L}.csv python tools/profile_frame.py --workload $wl --layout $L --frames 3 \
      > gpurun_out/s1_ncu_${wl}_${L}.log 2>&1
  done
done
for L in rowmajor tile4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_render_lin -s 1 -c 1 \
    -o gpurun_out/s1_full_flythrough4k_${L} python tools/profile_frame.py --workload flythrough4k --layout $L --frames 3 \
    > gpurun_out/s1_full_${L}.log 2>&1
done
# K1: every launch of the prepass at 16384^2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_ -c 40 --csv \
  --log-file gpurun_out/s1_k1_launches.csv python tools/profile_frame.py --workload flythrough4k --frames 1 > gpurun_out/s1_k1.log 2>&1
ls -la gpurun_out | tail -40
