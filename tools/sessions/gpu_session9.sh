#!/bin/bash
# GPU session 9: warp-level count kernel, rasterised GEMM tiles, unrolled Welford: parity, then timings
mkdir -p gpurun_out
timeout 300 python tools/dense_check.py > gpurun_out/s9_dense_check.txt 2>&1; echo "dense_check rc=$?"
grep -E "differ|rc=" gpurun_out/s9_dense_check.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s9_pytest.txt 2>&1
tail -4 gpurun_out/s9_pytest.txt
timeout 600 python tools/run_configs.py > gpurun_out/s9_configs.jsonl 2> gpurun_out/s9_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/s9_configs.jsonl'):
    d=json.loads(l); print(d['config'], d['host_s'], d['device_ms'])
PY
rm -f gpurun_out/s9_steps.txt
for opts in "--n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" "--n 20000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" \
            "--n 20000 --len 100 --g 12 --m 6 --batch 24 --acc-path 3"; do
  echo "== $opts" >> gpurun_out/s9_steps.txt
  timeout 300 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s9_steps.txt
done
cat gpurun_out/s9_steps.txt
