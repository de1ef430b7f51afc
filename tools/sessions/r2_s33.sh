#!/bin/bash
# r2 session 33 (1 GPU): unequal output shares on one GPU (parity of the sharded finalisation)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "shards_on_one_gpu or golden or api_surface" > gpurun_out/r2s33_pytest.txt 2>&1
tail -4 gpurun_out/r2s33_pytest.txt
