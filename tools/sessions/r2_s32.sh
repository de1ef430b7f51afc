#!/bin/bash
# r2 session 32 (1 GPU): window position carried through the sort (no atomic in the task filing): parity, then C4 steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window_position or directory_form or random_exact or heavy or skewed or golden or 32bit" > gpurun_out/r2s32_pytest.txt 2>&1
tail -4 gpurun_out/r2s32_pytest.txt
NCOMB=768 timeout 900 python tools/c4_steps.py '{"count_updates": 0, "seg_side": 1}' '{"count_updates": 0, "seg_side": 2}' > gpurun_out/r2s32_steps.txt 2>&1
cat gpurun_out/r2s32_steps.txt
