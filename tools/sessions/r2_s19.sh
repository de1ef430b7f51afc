#!/bin/bash
# r2 session 19 (8 GPUs): where the fixed costs of an 8-rank job go (phases per rank), after the two-pass normalisation and the IPC mapping cache
mkdir -p gpurun_out
FSK_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_ranks.py \
    > gpurun_out/r2s19_e2e.txt 2> gpurun_out/r2s19_e2e.err
grep -v "^\[" gpurun_out/r2s19_e2e.txt | tail -6
grep "rank 0 \|rank 7 " gpurun_out/r2s19_e2e.txt | tail -14
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
