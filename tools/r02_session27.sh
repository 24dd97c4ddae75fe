#!/bin/bash
# Round-2 GPU session 27 (1 GPU): ncu counters of ONE band launch (rank 0's interleaved tile rows of an 8K frame split
# 2 / 4 / 8 ways) for the roofline of the bands8k sub-record at N > 1.  Each capture after the same command ran plain.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for n in 8 4 2; do
  python tools/profile_frame.py --workload bands8k --frames 3 --rgb8 --band $n:0 > gpurun_out/s27_plain_$n.log 2>&1 &&
  timeout 300 ncu --metrics $M --clock-control none -k regex:k2_render_lin -s 1 -c 2 --csv \
    --log-file gpurun_out/s27_ncu_bands8kx$n.csv python tools/profile_frame.py --workload bands8k --frames 3 --rgb8 --band $n:0 > gpurun_out/s27_ncu_$n.log 2>&1
  echo "n=$n exit $?"; tail -n 1 gpurun_out/s27_plain_$n.log
done
