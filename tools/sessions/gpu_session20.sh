#!/bin/bash
# GPU session 20: whole parity suite with the heavy-run stage, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s20_pytest.txt 2>&1
tail -6 gpurun_out/s20_pytest.txt
timeout 900 python bench.py > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s20_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_sort']['frac'], d['gpu_launches'])
PY
tail -2 gpurun_out/s20_bench.err
