#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2_pytest.txt 2>&1
tail -15 gpurun_out/s2_pytest.txt
for opts in "--pad 1" "--pad 2" "--pad 2 --rows-threads 768" "--pad 2 --rows-threads 512" "--pad 2 --ld-hint 1" "--pad 2 --ld-hint 2"; do
  echo "== $opts" >> gpurun_out/s2_steps.txt
  timeout 300 python tools/profile_step.py --batch 48 --reps 2 $opts >> gpurun_out/s2_steps.txt 2>&1
done
cat gpurun_out/s2_steps.txt
