#!/bin/bash
# r2 session 23 (1 GPU): overlap of the pre-pass with the accumulate -- stream priorities, common shared-memory carve-out
mkdir -p gpurun_out
timeout 900 python tools/c4_steps.py '{"count_updates": 0, "overlap": 1}' '{"count_updates": 0, "overlap": 2}' '{"count_updates": 0, "overlap": 3}' '{"count_updates": 0, "overlap": 2, "wave": 2}' > gpurun_out/r2s23_steps.txt 2>&1
cat gpurun_out/r2s23_steps.txt
