#!/bin/bash
# GPU session 25: two-tile shape of the tcgen05 contraction (B operand shared by two accumulators): parity, then timings per shape
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "dense or heavy or approx" ) > gpurun_out/s25_pytest.txt 2>&1
grep -E "passed|failed|Error" gpurun_out/s25_pytest.txt | tail -3
timeout 300 python tools/dense_check.py 2>&1 | grep -E "differ"
rm -f gpurun_out/s25_steps.txt
for shape in 1 2; do
for opts in "--n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" "--n 20000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" "--skew 1 --batch 96 --reps 1"; do
  echo "== shape $shape $opts" >> gpurun_out/s25_steps.txt
  timeout 600 python tools/profile_step.py --reps 2 --gemm-shape $shape $opts 2>&1 | head -1 >> gpurun_out/s25_steps.txt
done; done
cut -c1-330 gpurun_out/s25_steps.txt
