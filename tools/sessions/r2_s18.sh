#!/bin/bash
# r2 session 18 (8 GPUs): multi-rank parity tests, the full build under torchrun at N=8, the same build from one plain process
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s18_pytest.txt 2>&1
tail -3 gpurun_out/r2s18_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 \
    > gpurun_out/r2s18_bench_n8.json 2> gpurun_out/r2s18_bench_n8.err
tail -c 600 gpurun_out/r2s18_bench_n8.json; tail -3 gpurun_out/r2s18_bench_n8.err
timeout 600 python tools/team_c4.py > gpurun_out/r2s18_team.txt 2> gpurun_out/r2s18_team.err
cat gpurun_out/r2s18_team.txt; tail -3 gpurun_out/r2s18_team.err
