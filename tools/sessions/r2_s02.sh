#!/bin/bash
# r2 session 2 (1 GPU): directory form of the segmentation -- parity tests, then C4 steps with seg_dir off / on / 64 blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "directory_form or skewed or heavy" > gpurun_out/r2s02_pytest.txt 2>&1
tail -5 gpurun_out/r2s02_pytest.txt
timeout 900 python tools/c4_steps.py > gpurun_out/r2s02_steps.txt 2>&1
cat gpurun_out/r2s02_steps.txt
