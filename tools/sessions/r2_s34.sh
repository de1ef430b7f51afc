#!/bin/bash
# r2 session 34 (8 GPUs): results handed back through the GPUs with the faster path to host memory (probed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s34_pytest.txt 2>&1
tail -3 gpurun_out/r2s34_pytest.txt
FSK_TRACE=1 timeout 600 python tools/team_c4.py > gpurun_out/r2s34_team.txt 2> gpurun_out/r2s34_team.err
cat gpurun_out/r2s34_team.txt; grep "device -> host" gpurun_out/r2s34_team.err | head -3
FSK_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_ranks.py \
    > gpurun_out/r2s34_e2e.txt 2> gpurun_out/r2s34_e2e.err
grep -v "^\[" gpurun_out/r2s34_e2e.txt | grep "rep\|parity" | tail -5; grep "device -> host" gpurun_out/r2s34_e2e.txt | head -2
grep "rank 4 finalize\|rank 0 finalize" gpurun_out/r2s34_e2e.txt | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 \
    > gpurun_out/r2s34_bench_n8.json 2> gpurun_out/r2s34_bench_n8.err
tail -2 gpurun_out/r2s34_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s34_bench_n8.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "wall_s_per_build", "parity_ok")}, d["e2e"]["wall_s"], d["e2e"]["parts_rank0"])
PY
