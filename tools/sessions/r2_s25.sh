#!/bin/bash
# r2 session 25 (1 GPU): where the time of the AImed / protein configurations goes
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r2s25_configs.txt 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.aimed_workload(0)))
print(json.dumps(bench.fasta_workload("1.1", 10, 6, 0)))
print(json.dumps(bench.fasta_workload("AImed", 20, 10, 0, t=20, approx=True, max_iters=50, skip_variance=True)))
PY
cat gpurun_out/r2s25_configs.txt
