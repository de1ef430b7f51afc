#!/bin/bash
# r2 session 27 (1 GPU): batches of 384 combinations (25 KB of kernel parameters) against 192
mkdir -p gpurun_out
NCOMB=768 timeout 900 python tools/c4_steps.py '{"count_updates": 0, "batch": 192}' '{"count_updates": 0, "batch": 384}' '{"count_updates": 0, "batch": 256}' > gpurun_out/r2s27_steps.txt 2>&1
cat gpurun_out/r2s27_steps.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_exact or golden or approx or dense or heavy or speculated" > gpurun_out/r2s27_pytest.txt 2>&1
tail -3 gpurun_out/r2s27_pytest.txt
