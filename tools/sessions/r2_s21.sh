#!/bin/bash
# r2 session 21 (1 GPU): what the driver runs at round end -- GPU suite, smoke, reference arm, the bench with the driver's flags
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s21_pytest.txt 2>&1
tail -6 gpurun_out/r2s21_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s21_ref.json 2> gpurun_out/r2s21_ref.err
tail -c 400 gpurun_out/r2s21_ref.json; tail -4 gpurun_out/r2s21_ref.err
( time timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s21_bench.json 2> gpurun_out/r2s21_bench.err
tail -c 300 gpurun_out/r2s21_bench.json; tail -4 gpurun_out/r2s21_bench.err
