#!/bin/bash
# GPU session 16: column windows of the row path; whole parity suite; default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s16_pytest.txt 2>&1
tail -5 gpurun_out/s16_pytest.txt
timeout 300 python tools/profile_step.py --batch 96 --reps 2 2>&1 | head -1
timeout 300 python tools/profile_step.py --n 60000 --batch 48 --reps 1 2>&1 | head -2
timeout 900 python bench.py > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s16_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_sort']['frac'], d['other_workloads']['dense_tensor_core']['device_ms'])
PY
