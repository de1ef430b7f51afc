#!/bin/bash
# r2 session 30 (1 GPU): final state -- GPU suite, smoke, example scripts, default bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s30_pytest.txt 2>&1
tail -5 gpurun_out/r2s30_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python examples/run_check.py 2>&1 | tail -2
timeout 300 python examples/thread_time.py --dataset EP300 -g 10 -m 6 --output-dir gpurun_out 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2s30_bench.json 2> gpurun_out/r2s30_bench.err
tail -3 gpurun_out/r2s30_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s30_bench.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "wall_s_per_build", "parity_ok", "gpu_launches")}, d["e2e"]["wall_s"], d["config"]["batch"])
print(d["roofline"]["frac"], d["roofline"]["frac_measured_traffic"], d["roofline_sort"]["frac"], d["roofline_sort"]["per_stage_gbs"])
print(json.dumps(d["other_workloads"]["dense_tensor_core"])[:300])
print(json.dumps(d["other_workloads"]["ep300_approx_t1"])[:260])
print(json.dumps(d["other_workloads"]["skewed"])[:600])
PY
