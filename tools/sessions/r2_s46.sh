#!/bin/bash
# r2 session 46 (1 GPU): dense_count_kernel holds the windows of short sequences in registers -- dense tests, EP300 exact / approx timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or approx or variance or speculated or registers or golden or fingerprints or random_exact" > gpurun_out/r2s46_pytest.txt 2>&1
tail -3 gpurun_out/r2s46_pytest.txt
timeout 300 python - > gpurun_out/r2s46_ep300.txt 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for kw in (dict(), dict(approx=True, t=1, max_iters=50), dict(approx=True, t=20, max_iters=50)):
    best = None
    for rep in range(5):
        f = FastSK(10, 6, seed=0, device=0, distributed=False, profile=True, **kw)
        f.compute_kernel(Xtr, Xte); st = f.stats()
        row = {"cfg": kw, "device_ms": round(st["ms_total"], 3), "combos": st["combos_done"], "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")}}
        if best is None or row["device_ms"] < best["device_ms"]: best = row
    print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s46_ep300.txt
