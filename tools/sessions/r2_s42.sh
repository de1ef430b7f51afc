#!/bin/bash
# r2 session 42 (2 GPUs): the multi-rank tests after the dense-regime changes (register-form Welford, byte operands)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s42_pytest_multi.txt 2>&1
tail -4 gpurun_out/r2s42_pytest_multi.txt
