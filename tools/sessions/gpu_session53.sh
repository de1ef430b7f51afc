#!/bin/bash
# GPU session 53 (4 GPUs): the driver's multi-rank launch of the bench at N = 4
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 4 --warmup 3 \
    > gpurun_out/s53_bench_n4.json 2> gpurun_out/s53_bench_n4.err
tail -c 300 gpurun_out/s53_bench_n4.json; echo
