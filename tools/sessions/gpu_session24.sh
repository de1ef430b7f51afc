#!/bin/bash
# GPU session 24: batches of up to 192 combinations
mkdir -p gpurun_out
rm -f gpurun_out/s24_steps.txt
for opts in "--batch 96" "--batch 144" "--batch 192"; do
  echo "== $opts" >> gpurun_out/s24_steps.txt
  timeout 600 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s24_steps.txt
done
cat gpurun_out/s24_steps.txt | cut -c1-400
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s24_pytest.txt 2>&1
grep -E "passed|failed" gpurun_out/s24_pytest.txt | tail -1
