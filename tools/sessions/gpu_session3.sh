#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/s3_steps.txt
for opts in "--overlap 0" "--overlap 1" "--overlap 1 --rows-threads 768" "--overlap 1 --rows-threads 896" "--overlap 1 --wave 2" "--overlap 1 --wave 8"; do
  echo "== $opts" >> gpurun_out/s3_steps.txt
  timeout 300 python tools/profile_step.py --batch 48 --reps 2 --batches-per-call 4 $opts >> gpurun_out/s3_steps.txt 2>&1
done
cat gpurun_out/s3_steps.txt
