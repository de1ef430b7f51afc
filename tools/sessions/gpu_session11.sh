#!/bin/bash
# GPU session 11: fused last sort pass + segmentation (bucket_segment_kernel): parity, then C4 step timings against the two-pass path
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "fused or c4_shape" ) > gpurun_out/s11_pytest_fused.txt 2>&1
tail -12 gpurun_out/s11_pytest_fused.txt
rm -f gpurun_out/s11_steps.txt
for opts in "--batch 96 --seg-fused 1" "--batch 96 --seg-fused 2" "--batch 48 --seg-fused 2"; do
  echo "== $opts" >> gpurun_out/s11_steps.txt
  timeout 300 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s11_steps.txt
done
cat gpurun_out/s11_steps.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s11_pytest.txt 2>&1
tail -4 gpurun_out/s11_pytest.txt
