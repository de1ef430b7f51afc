#!/bin/bash
# GPU session 26: after pruning the experiment knobs: parity suite, C4 step, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s26_pytest.txt 2>&1
grep -E "passed|failed" gpurun_out/s26_pytest.txt | tail -1
timeout 300 python tools/profile_step.py --batch 192 --reps 2 2>&1 | head -1 | cut -c1-330
timeout 900 python bench.py > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s26_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_sort']['frac'], d['gpu_launches'], d['config']['batch'])
print(json.dumps(d['other_workloads'])[:1500])
PY
tail -2 gpurun_out/s26_bench.err
