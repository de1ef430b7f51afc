#!/bin/bash
# GPU session 21: heavy-run threshold sweep on the skewed set; bench with the skewed section
mkdir -p gpurun_out
rm -f gpurun_out/s21_steps.txt
for tau in 5000 2500 1500 1000 600; do
  echo "== --skew 1 --batch 96 --heavy-tau $tau" >> gpurun_out/s21_steps.txt
  timeout 600 python tools/profile_step.py --reps 1 --skew 1 --batch 96 --heavy-tau $tau 2>&1 | head -1 >> gpurun_out/s21_steps.txt
done
cat gpurun_out/s21_steps.txt | cut -c1-400
timeout 900 python bench.py --steps 4 --no-cpu-baseline > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s21_bench.json'))
print(d['value'], json.dumps(d['other_workloads']['skewed']))
PY
tail -2 gpurun_out/s21_bench.err
