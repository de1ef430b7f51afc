#!/bin/bash
# r2 session 24 (1 GPU): ncu launch list of the bench command (kernel shares of the step), ncu --set full of the final accumulate
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2s24_launches.csv \
    python bench.py --steps 1 --warmup 3 --combos 768 --no-parity --no-cpu-baseline --no-skewed > gpurun_out/r2s24_bench_under_ncu.json 2> gpurun_out/r2s24_bench_under_ncu.err
wc -l gpurun_out/r2s24_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_rows -c 1 -f -o gpurun_out/r2s24_acc_final \
      python tools/c4_steps.py '{"batch": 48, "wave": 400, "count_updates": 0}' > gpurun_out/r2s24_acc.log 2>&1
tail -2 gpurun_out/r2s24_acc.log
