#!/bin/bash
# r2 session 22 (1 GPU): pre-pass of the next batch overlapped with the accumulate (48-register row kernel + 25 KB sort CTAs)
mkdir -p gpurun_out
timeout 900 python tools/c4_steps.py '{"count_updates": 0}' '{"count_updates": 0, "overlap": 1}' '{"count_updates": 0, "overlap": 1, "wave": 8}' '{"count_updates": 0, "overlap": 1, "fit_smem": 0}' > gpurun_out/r2s22_steps.txt 2>&1
cat gpurun_out/r2s22_steps.txt
