#!/bin/bash
# r2 session 43 (1 GPU): byte columns for the heavy-run contraction -- heavy / dense tests, skewed set by threshold
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "heavy or dense or directory or registers" > gpurun_out/r2s43_pytest.txt 2>&1
tail -3 gpurun_out/r2s43_pytest.txt
timeout 600 python tools/skewed_steps.py '{}' '{"dense_u8": 0}' '{"heavy_tau": 2000}' '{"heavy_tau": 1650}' '{"heavy_tau": 1300}' '{"gemm_shape": 1}' > gpurun_out/r2s43_skewed.txt 2>&1
cut -c1-330 gpurun_out/r2s43_skewed.txt
