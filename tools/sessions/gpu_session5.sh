#!/bin/bash
# GPU session 5: parity with the rotate-and-mask pack and batch 96; batch size and segment occupancy variants
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s5_pytest.txt 2>&1
tail -3 gpurun_out/s5_pytest.txt
rm -f gpurun_out/s5_steps.txt
for opts in "--batch 48" "--batch 96" "--batch 64" "--batch 48 --seg-occ 1" "--batch 48 --seg-occ 2" "--batch 48 --seg-occ 3" "--batch 96 --seg-occ 2 --wave 8" "--batch 96 --wave 2"; do
  echo "== $opts" >> gpurun_out/s5_steps.txt
  timeout 300 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s5_steps.txt
done
cat gpurun_out/s5_steps.txt
