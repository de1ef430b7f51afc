#!/bin/bash
# GPU session 15: evidence for the current pipeline: launch list of the bench command, full ncu capture of the accumulate at
# batch 96 (one launch over all rows), of pack/sort/segment, and of the dense kernels after the epilogue change
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s15_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s15_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/s15_bench_under_ncu.log; echo
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"accumulate_rows" -c 1 -o gpurun_out/s15_acc96 \
    python tools/profile_step.py --batch 96 --wave 400 --reps 0 > gpurun_out/s15_ncu_acc.log 2>&1
tail -2 gpurun_out/s15_ncu_acc.log
timeout 900 ncu --set full --clock-control none -k regex:"pack_hist|onesweep|segment_kernel" -c 4 -o gpurun_out/s15_presort96 \
    python tools/profile_step.py --batch 96 --reps 0 > gpurun_out/s15_ncu_pre.log 2>&1
tail -2 gpurun_out/s15_ncu_pre.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"syrk_tc|dense_count" -s 2 -c 2 -o gpurun_out/s15_dense \
    python tools/profile_step.py --n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3 --reps 1 > gpurun_out/s15_ncu_dense.log 2>&1
tail -2 gpurun_out/s15_ncu_dense.log
