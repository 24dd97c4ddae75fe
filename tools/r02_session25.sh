#!/bin/bash
# Round-2 GPU session 25 (1 GPU): ncu --set full of one k2_render_lin launch on the FINAL build (flythrough4k frame 1,
# RGB8), after the same command has exited 0 without ncu.
mkdir -p gpurun_out
python tools/profile_frame.py --workload flythrough4k --frames 3 --rgb8 > gpurun_out/s25_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_render_lin -s 1 -c 1 \
  -o gpurun_out/s25_full_flythrough4k python tools/profile_frame.py --workload flythrough4k --frames 3 --rgb8 > gpurun_out/s25_full.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/s25_plain.log
