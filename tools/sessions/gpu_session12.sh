#!/bin/bash
# GPU session 12: ncu capture of bucket_segment_kernel on the C4 workload (batch 8)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"bucket_segment" -c 1 -o gpurun_out/s12_bucket \
    python tools/profile_step.py --batch 8 --seg-fused 2 --reps 0 > gpurun_out/s12_ncu.log 2>&1
tail -2 gpurun_out/s12_ncu.log
