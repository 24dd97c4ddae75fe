#!/bin/bash
# Round-2 GPU session 7: stack-less k2_render_lin + shuffle dilation — tests, timings, K1 launch list, counters.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/s7_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s7_pytest.log
tail -3 gpurun_out/s7_pytest.log
for wl in flythrough4k ortho4k spherical1080 bands8k sample720; do
  echo "== $wl skip"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 10
  echo "== $wl skip_rgb8"
  timeout 300 python tools/profile_frame.py --workload $wl --frames 10 --rgb8
done > gpurun_out/s7_variants.txt 2>&1
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active
for wl in flythrough4k sample720 spherical1080 ortho4k bands8k; do
  python tools/profile_frame.py --workload $wl --frames 3 --rgb8 > gpurun_out/s7_plain_$wl.log 2>&1 &&
  timeout 600 ncu --metrics $M --clock-control none -k regex:k2_render_lin -s 1 -c 2 --csv \
    --log-file gpurun_out/s7_ncu_$wl.csv python tools/profile_frame.py --workload $wl --frames 3 --rgb8 > gpurun_out/s7_ncu_$wl.log 2>&1
done
python tools/profile_frame.py --workload flythrough4k --frames 3 --rgb8 > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_render_lin -s 1 -c 1 \
  -o gpurun_out/s7_full_flythrough4k python tools/profile_frame.py --workload flythrough4k --frames 3 --rgb8 > gpurun_out/s7_full.log 2>&1
python tools/profile_frame.py --workload flythrough4k --frames 1 > /dev/null 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k1_ -c 40 --csv \
  --log-file gpurun_out/s7_k1_launches.csv python tools/profile_frame.py --workload flythrough4k --frames 1 > gpurun_out/s7_k1.log 2>&1
