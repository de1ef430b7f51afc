#!/bin/bash
# r2 session 20 (8 GPUs): the bench at N=8 after the two-pass normalisation / IPC cache / interleaved host buffers; the one-process team
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 \
    > gpurun_out/r2s20_bench_n8.json 2> gpurun_out/r2s20_bench_n8.err
tail -c 300 gpurun_out/r2s20_bench_n8.json; tail -3 gpurun_out/r2s20_bench_n8.err
timeout 600 python tools/team_c4.py > gpurun_out/r2s20_team.txt 2> gpurun_out/r2s20_team.err
cat gpurun_out/r2s20_team.txt; tail -3 gpurun_out/r2s20_team.err
