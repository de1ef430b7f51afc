#!/bin/bash
# GPU session 14: Welford step fused into the accumulate flush (row path) and the GEMM epilogue (dense path)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s14_pytest.txt 2>&1
tail -6 gpurun_out/s14_pytest.txt
timeout 600 python tools/run_configs.py > gpurun_out/s14_configs.jsonl 2> gpurun_out/s14_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/s14_configs.jsonl'):
    d=json.loads(l); print(d['config'], d['host_s']['build_partial'], d['host_s']['total'], d['device_ms'])
PY
tail -3 gpurun_out/s14_configs.err
