#!/bin/bash
# GPU session 8: dense tensor-core path against the row path on the bundled configurations (device + host phases),
# EP300-shaped synthetic inputs at N = 4000 and 20000, and a full ncu capture of the two dense kernels
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "dense or approx" ) > gpurun_out/s8_pytest_dense.txt 2>&1
tail -4 gpurun_out/s8_pytest_dense.txt
timeout 600 python tools/run_configs.py > gpurun_out/s8_configs.jsonl 2> gpurun_out/s8_configs.err
cut -c1-700 gpurun_out/s8_configs.jsonl
rm -f gpurun_out/s8_steps.txt
for opts in "--n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 2" "--n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" \
            "--n 20000 --len 100 --g 10 --m 6 --batch 96 --acc-path 2" "--n 20000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" \
            "--n 20000 --len 100 --g 12 --m 6 --batch 96 --acc-path 2" "--n 20000 --len 100 --g 12 --m 6 --batch 24 --acc-path 3"; do
  echo "== $opts" >> gpurun_out/s8_steps.txt
  timeout 300 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s8_steps.txt
done
cat gpurun_out/s8_steps.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"syrk_tc|dense_count" -c 2 -o gpurun_out/s8_dense \
    python tools/profile_step.py --n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3 --reps 0 > gpurun_out/s8_ncu.log 2>&1
tail -2 gpurun_out/s8_ncu.log
