#!/bin/bash
# r2 session 44 (1 GPU): bench.py with every section (parity, cpu baseline, skewed, other workloads) on a shortened build + smoke()
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s44_smoke.txt 2>&1; tail -2 gpurun_out/r2s44_smoke.txt
timeout 600 python bench.py --steps 1 --warmup 3 --combos 768 > gpurun_out/r2s44_bench.json 2> gpurun_out/r2s44_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s44_bench.json").read().strip().splitlines()[-1])
    o = d["other_workloads"]
    print("value", d["value"], "parity_ok", d["parity_ok"], "e2e", d["e2e"]["wall_s"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
    print("skewed", o.get("skewed", {}).get("rows_plus_heavy_runs_on_tensor_cores"))
    print("same_config", o.get("same_config"))
    for k in ("dense_tensor_core", "ep300_approx_t1", "protein_1_1", "aimed_approx"):
        print(k, {a: o[k][a] for a in o[k] if a in ("e2e_s", "device_ms", "ms_accumulate", "tensor_tflops", "combinations_per_s_device")})
except Exception as e:
    print("bench:", e); print(open("gpurun_out/r2s44_bench.err").read()[-1500:])
PY
