#!/bin/bash
# r2 session 9 (1 GPU): whole GPU suite (reference fingerprints of the full EP300 / protein sets), EP300 approx timing,
# host-side phases after the fixes, the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s09_pytest.txt 2>&1
tail -4 gpurun_out/r2s09_pytest.txt
timeout 300 python tools/host_profile.py ep300 > gpurun_out/r2s09_host_ep300.txt 2>&1
tail -22 gpurun_out/r2s09_host_ep300.txt
timeout 900 python bench.py > gpurun_out/r2s09_bench.json 2> gpurun_out/r2s09_bench.err
tail -3 gpurun_out/r2s09_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s09_bench.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "wall_s_per_build", "parity_ok", "gpu_launches", "phase_ms_per_step")})
print(d["e2e"]["wall_s"], d["roofline"]["frac"], d["roofline_sort"]["frac"], d["roofline_sort"]["per_stage_gbs"])
ow = d["other_workloads"]
for k in ("ep300_approx_t1", "dense_tensor_core", "protein_1_1", "aimed_approx", "same_config", "skewed"):
    print(k, json.dumps(ow.get(k))[:600])
PY
