#!/bin/bash
# GPU session 48: ncu captures of the final segment kernel (12 rows per warp) and of the two-tile contraction (heavy runs, skewed set)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"segment_kernel" -c 1 -o gpurun_out/s48_segment \
    python tools/profile_step.py --batch 96 --reps 0 > gpurun_out/s48_ncu_seg.log 2>&1
tail -1 gpurun_out/s48_ncu_seg.log
timeout 600 ncu --set full --clock-control none -k regex:"syrk_tc|heavy_fill|heavy_zero" -c 3 -o gpurun_out/s48_heavy \
    python tools/profile_step.py --skew 1 --batch 24 --reps 0 > gpurun_out/s48_ncu_heavy.log 2>&1
tail -1 gpurun_out/s48_ncu_heavy.log
