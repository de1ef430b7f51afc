#!/bin/bash
# r2 session 29 (8 GPUs): final state -- multi-rank parity tests, bench at N=8 and N=4, the one-process team, GPU-count sweep on EP300
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s29_pytest.txt 2>&1
tail -3 gpurun_out/r2s29_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 \
    > gpurun_out/r2s29_bench_n8.json 2> gpurun_out/r2s29_bench_n8.err
tail -c 200 gpurun_out/r2s29_bench_n8.json; tail -2 gpurun_out/r2s29_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 2 --warmup 3 \
    > gpurun_out/r2s29_bench_n4.json 2> gpurun_out/r2s29_bench_n4.err
tail -c 200 gpurun_out/r2s29_bench_n4.json; tail -2 gpurun_out/r2s29_bench_n4.err
timeout 600 python tools/team_c4.py > gpurun_out/r2s29_team.txt 2> gpurun_out/r2s29_team.err
cat gpurun_out/r2s29_team.txt; tail -2 gpurun_out/r2s29_team.err
