#!/bin/bash
# GPU session 1: re-validate parity, baseline bench on this pod, L2 fetch-granularity experiments on the accumulate
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_pytest.txt 2>&1
for opts in "" "--l2-fetch 32" "--l2-fetch 128" "--ld-hint 1" "--ld-hint 2" "--ld-hint 1 --l2-fetch 32"; do
  echo "== $opts" >> gpurun_out/s1_steps.txt
  python tools/profile_step.py --batch 48 --reps 2 $opts >> gpurun_out/s1_steps.txt 2>&1
done
for opts in "" "--l2-fetch 32" "--ld-hint 1" "--ld-hint 1 --l2-fetch 32"; do
  tag=$(echo "x$opts" | tr -d ' -')
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:accumulate_rows --csv \
      --log-file gpurun_out/s1_ncu_$tag.csv python tools/profile_step.py --batch 8 --wave 400 --reps 1 $opts > gpurun_out/s1_ncu_$tag.log 2>&1
done
python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_pytest.txt; cat gpurun_out/s1_steps.txt
