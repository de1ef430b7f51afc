#!/bin/bash
# r2 session 5 (1 GPU): register-blocked segmentation kernel -- parity (whole GPU suite), C4 steps in every form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2s05_pytest.txt 2>&1
tail -4 gpurun_out/r2s05_pytest.txt
timeout 900 python tools/c4_steps.py '{"seg_lean": 1}' '{"seg_lean": 2}' '{"seg_lean": 2, "count_updates": 0}' '{"seg_lean": 2, "seg_dir": 2, "count_updates": 0}' '{"seg_lean": 2, "pad": 1, "count_updates": 0}' > gpurun_out/r2s05_steps.txt 2>&1
cat gpurun_out/r2s05_steps.txt
