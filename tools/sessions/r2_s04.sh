#!/bin/bash
# r2 session 4 (1 GPU): three-deep chunk pipeline of the accumulate (both task forms), host-side phase timing
mkdir -p gpurun_out
timeout 900 python tools/c4_steps.py '{"seg_dir": 1}' '{"seg_dir": 2}' '{"seg_dir": 1, "acc_unroll": 4}' '{"seg_dir": 2, "acc_unroll": 4}' '{"seg_dir": 1, "acc_prefetch": 0}' > gpurun_out/r2s04_steps.txt 2>&1
cat gpurun_out/r2s04_steps.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2s04_pytest.txt 2>&1
tail -3 gpurun_out/r2s04_pytest.txt
timeout 300 python tools/host_profile.py ep300 > gpurun_out/r2s04_host_ep300.txt 2>&1
tail -32 gpurun_out/r2s04_host_ep300.txt
timeout 300 python tools/host_profile.py c4 384 > gpurun_out/r2s04_host_c4.txt 2>&1
tail -24 gpurun_out/r2s04_host_c4.txt
