#!/bin/bash
# r2 session 6 (1 GPU): register-blocked segmentation, task form against directory form; ncu of the lean kernel
mkdir -p gpurun_out
timeout 900 python tools/c4_steps.py '{"seg_lean": 2, "seg_dir": 1, "count_updates": 0}' '{"seg_lean": 2, "seg_dir": 2, "count_updates": 0}' '{"seg_lean": 1, "seg_dir": 1, "count_updates": 0}' > gpurun_out/r2s06_steps.txt 2>&1
cat gpurun_out/r2s06_steps.txt
for d in 1 2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:segment_lean -c 1 -f -o gpurun_out/r2s06_lean_dir$d \
      python tools/c4_steps.py "{\"seg_dir\": $d, \"seg_lean\": 2, \"batch\": 48, \"count_updates\": 0}" > gpurun_out/r2s06_lean_dir$d.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "directory_form" > gpurun_out/r2s06_pytest.txt 2>&1
tail -4 gpurun_out/r2s06_pytest.txt
