#!/bin/bash
# r2 session 7 (2 GPUs): multi-rank parity tests (team in one process, torchrun peer / all-reduce), bench at N=2, team C4 job
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s07_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s07_pytest.txt 2>&1
tail -15 gpurun_out/r2s07_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 \
    > gpurun_out/r2s07_bench_n2.json 2> gpurun_out/r2s07_bench_n2.err
tail -c 1200 gpurun_out/r2s07_bench_n2.json; tail -5 gpurun_out/r2s07_bench_n2.err
FSK_TRACE=1 timeout 600 python tools/team_c4.py > gpurun_out/r2s07_team.txt 2> gpurun_out/r2s07_team.err
cat gpurun_out/r2s07_team.txt; tail -30 gpurun_out/r2s07_team.err
