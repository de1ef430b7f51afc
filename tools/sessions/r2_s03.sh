#!/bin/bash
# r2 session 3 (1 GPU): ncu --set full of segment_kernel and accumulate_rows_kernel, task-list form against directory form
mkdir -p gpurun_out
for mode in 1 2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:segment_kernel -c 1 -f -o gpurun_out/r2s03_seg_dir$mode \
      python tools/c4_steps.py "{\"seg_dir\": $mode, \"batch\": 48}" > gpurun_out/r2s03_seg_dir$mode.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_rows -c 1 -f -o gpurun_out/r2s03_acc_dir$mode \
      python tools/c4_steps.py "{\"seg_dir\": $mode, \"batch\": 48, \"wave\": 400}" > gpurun_out/r2s03_acc_dir$mode.log 2>&1
done
ls -la gpurun_out/r2s03*
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "directory_form" > gpurun_out/r2s03_pytest.txt 2>&1
tail -3 gpurun_out/r2s03_pytest.txt
