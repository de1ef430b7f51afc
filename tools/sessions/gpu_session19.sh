#!/bin/bash
# GPU session 19: heavy runs on the tensor cores: parity, then the skewed full-size workload with and without, and the uniform one
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "heavy or skewed" ) > gpurun_out/s19_pytest_heavy.txt 2>&1
tail -12 gpurun_out/s19_pytest_heavy.txt
rm -f gpurun_out/s19_steps.txt
for opts in "--skew 1 --batch 96 --heavy-tau -1" "--skew 1 --batch 96" "--batch 96 --heavy-tau -1 --reps 2" "--batch 96 --reps 2"; do
  echo "== $opts" >> gpurun_out/s19_steps.txt
  timeout 600 python tools/profile_step.py --reps 1 $opts 2>&1 | head -1 >> gpurun_out/s19_steps.txt
done
cat gpurun_out/s19_steps.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s19_pytest.txt 2>&1
tail -4 gpurun_out/s19_pytest.txt
