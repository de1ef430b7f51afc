#!/bin/bash
# r2 session 51 (1 GPU): bench.py after the tensor roofline object was added to other_workloads (shortened build)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --combos 384 --no-parity --no-cpu-baseline --no-skewed > gpurun_out/r2s51_bench_short.json 2> gpurun_out/r2s51_bench_short.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s51_bench_short.json").read().strip().splitlines()[-1])
    print("value", d["value"], d["other_workloads"]["dense_tensor_core"]["roofline"])
except Exception as e:
    print("bench:", e); print(open("gpurun_out/r2s51_bench_short.err").read()[-1500:])
PY
