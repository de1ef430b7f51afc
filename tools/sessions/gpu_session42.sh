#!/bin/bash
# GPU session 42: final state of round 1: default bench, launch list of the bench command, bundled configurations
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s42_bench.json 2> gpurun_out/s42_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s42_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['roofline_sort']['frac'], d['gpu_launches'], d['config']['batch'])
print(d['roofline_sort']['per_stage_gbs'], d['clocks'])
PY
tail -2 gpurun_out/s42_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s42_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-skewed > gpurun_out/s42_bench_under_ncu.log 2>&1
timeout 600 python tools/run_configs.py > gpurun_out/s42_configs.jsonl 2> gpurun_out/s42_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/s42_configs.jsonl'):
    d=json.loads(l); print(d['config'], d['host_s']['total'], d['device_ms']['total'])
PY
