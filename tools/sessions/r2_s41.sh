#!/bin/bash
# r2 session 41 (1 GPU): full GPU suite after the byte-operand / one-tile-default changes + the other_workloads section of the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s41_pytest.txt 2>&1
tail -4 gpurun_out/r2s41_pytest.txt
timeout 300 python bench.py --steps 1 --warmup 3 --combos 384 --no-parity --no-cpu-baseline --no-skewed > gpurun_out/r2s41_bench_short.json 2> gpurun_out/r2s41_bench_short.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s41_bench_short.json").read().strip().splitlines()[-1])
    o = d["other_workloads"]
    print("value", d["value"])
    for k in ("rows", "dense_tensor_core", "ep300_approx_t1"):
        print(k, {a: o[k][a] for a in o[k] if a in ("e2e_s", "device_ms", "ms_accumulate", "tensor_tflops", "combinations_per_s_device", "phase_ms")})
except Exception as e:
    print("bench:", e)
PY
