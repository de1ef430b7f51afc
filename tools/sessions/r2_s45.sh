#!/bin/bash
# r2 session 45 (1 GPU): the driver's bench command on the final code (default steps / warm-up)
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 > gpurun_out/r2s45_bench_n1.json 2> gpurun_out/r2s45_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s45_bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "parity_ok", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]["value"], d["e2e"]["wall_s"], "roofline", d["roofline"]["frac"], d["roofline_sort"]["frac"], "cpu", d["cpu_baseline"])
    print("phase", {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()})
except Exception as e:
    print("bench:", e); print(open("gpurun_out/r2s45_bench_n1.err").read()[-1500:])
PY
