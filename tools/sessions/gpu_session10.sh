#!/bin/bash
# GPU session 10: count-kernel fix, bench with the EP300 section
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "dense or approx or golden" ) > gpurun_out/s10_pytest.txt 2>&1
tail -4 gpurun_out/s10_pytest.txt
for opts in "--n 4000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3" "--n 20000 --len 100 --g 10 --m 6 --batch 96 --acc-path 3"; do
  timeout 300 python tools/profile_step.py --reps 2 $opts 2>&1 | head -1
done
timeout 900 python bench.py --steps 4 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['seconds'], d['e2e']['parts_rank0'])
print(json.dumps(d['other_workloads'], indent=1))
PY
tail -3 gpurun_out/s10_bench.err
