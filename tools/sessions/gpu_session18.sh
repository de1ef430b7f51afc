#!/bin/bash
# GPU session 18: skewed (Markov + planted motif) DNA: parity on both segmentation paths, full-size timing
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "skewed or column_windows" ) > gpurun_out/s18_pytest.txt 2>&1
tail -5 gpurun_out/s18_pytest.txt
rm -f gpurun_out/s18_steps.txt
for opts in "--skew 1 --batch 24 --seg-fused 1" "--skew 1 --batch 24 --seg-fused 2" "--skew 1 --batch 96"; do
  echo "== $opts" >> gpurun_out/s18_steps.txt
  timeout 600 python tools/profile_step.py --reps 1 $opts 2>&1 | head -1 >> gpurun_out/s18_steps.txt
done
cat gpurun_out/s18_steps.txt
