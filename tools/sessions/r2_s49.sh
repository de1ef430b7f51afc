#!/bin/bash
# r2 session 49 (2 GPUs): the driver's 2-GPU bench command on the final code (shortened build)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --combos 1536 > gpurun_out/r2s49_bench_n2.json 2> gpurun_out/r2s49_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s49_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "parity_ok")}, "e2e", d["e2e"]["wall_s"], d["e2e"]["parts_rank0"], d["step_parts_ms"]["merge_and_normalise"])
    print(d["other_workloads"].get("aimed_approx", {}).get("device_ms"))
except Exception as e:
    print("bench:", e); print(open("gpurun_out/r2s49_bench_n2.err").read()[-2000:])
PY
grep -i "device -> host\|weights" gpurun_out/r2s49_bench_n2.err | head -3
