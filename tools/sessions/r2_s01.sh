#!/bin/bash
# r2 session 1 (1 GPU): environment probes, full GPU test-suite, smoke, first full-build bench
mkdir -p gpurun_out
{ nvidia-smi -L; df -h /dev/shm; free -g | head -2; nproc; } > gpurun_out/r2s01_env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s01_pytest.txt 2>&1
tail -5 gpurun_out/r2s01_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s01_smoke.txt 2>&1
tail -2 gpurun_out/r2s01_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2s01_bench.json 2> gpurun_out/r2s01_bench.err
tail -c 600 gpurun_out/r2s01_bench.json; tail -5 gpurun_out/r2s01_bench.err
cat gpurun_out/r2s01_env.txt
