#!/bin/bash
# GPU session 22: adaptive heavy-run threshold: a list that is too small (4096 columns) must steer the threshold up within a few batches
mkdir -p gpurun_out
rm -f gpurun_out/s22_steps.txt
for opts in "--skew 1 --batch 96 --reps 1" "--skew 1 --batch 24 --heavy-cap 4096 --reps 1" "--skew 1 --batch 24 --heavy-cap 4096 --reps 3 --batches-per-call 4" "--skew 1 --batch 24 --reps 3 --batches-per-call 4"; do
  echo "== $opts" >> gpurun_out/s22_steps.txt
  timeout 600 python tools/profile_step.py $opts 2>&1 | head -1 >> gpurun_out/s22_steps.txt
done
cat gpurun_out/s22_steps.txt | cut -c1-400
( time timeout 600 python -m pytest tests -m gpu -x -q -k "heavy or skewed or c4" ) > gpurun_out/s22_pytest.txt 2>&1
tail -4 gpurun_out/s22_pytest.txt
