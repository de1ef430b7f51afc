#!/bin/bash
# r2 session 26 (1 GPU): where the time goes on the skewed set
mkdir -p gpurun_out
timeout 900 python tools/skewed_steps.py '{}' '{"heavy_tau": 1000}' '{"heavy_tau": 500}' > gpurun_out/r2s26_skewed.txt 2>&1
cat gpurun_out/r2s26_skewed.txt
