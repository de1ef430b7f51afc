#!/bin/bash
# GPU session 23: no per-batch fill of the id stream (segment pads the last unit of every run): parity + C4 step timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s23_pytest.txt 2>&1
tail -4 gpurun_out/s23_pytest.txt | head -2
timeout 300 python tools/profile_step.py --batch 96 --reps 2 2>&1 | head -1
timeout 300 python tools/profile_step.py --skew 1 --batch 96 --reps 1 2>&1 | head -1
