#!/bin/bash
# GPU session 28 (8 GPUs): the driver's multi-rank launch of the bench at N = 8
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 \
    > gpurun_out/s28_bench_n8.json 2> gpurun_out/s28_bench_n8.err
grep -c . gpurun_out/s28_bench_n8.json; tail -c 400 gpurun_out/s28_bench_n8.json; echo
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/s28_bench_n8.err | tail -5
