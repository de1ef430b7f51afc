#!/bin/bash
# r2 session 17 (1 GPU): row-per-thread Welford epilogue -- variance tests, EP300 t=1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "approx or variance or speculated or golden or column_windows" > gpurun_out/r2s17_pytest.txt 2>&1
tail -4 gpurun_out/r2s17_pytest.txt
timeout 300 python - > gpurun_out/r2s17_ep300_approx.txt 2>&1 <<'PY'
import json, sys, time
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for kw in (dict(t=1, max_iters=50), dict(t=20, max_iters=50), dict(t=1, max_iters=210)):
    best = None
    for rep in range(3):
        f = FastSK(10, 6, approx=True, seed=0, device=0, distributed=False, profile=True, **kw)
        t0 = time.perf_counter(); f.compute_kernel(Xtr, Xte); wall = time.perf_counter() - t0
        st = f.stats()
        row = {"cfg": kw, "wall_ms": round(wall * 1e3, 2), "device_ms": round(st["ms_total"], 3), "combos": st["combos_done"],
               "combos_per_s_device": round(st["combos_done"] / (st["ms_total"] * 1e-3)), "launches": st["kernel_launches"], "stdevs": len(f.get_stdevs()),
               "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")}}
        if best is None or row["device_ms"] < best["device_ms"]: best = row
    print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s17_ep300_approx.txt
timeout 300 python - <<'PY'
import sys, time, json
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for rep in range(3):
    f = FastSK(10, 6, seed=0, device=0, distributed=False, profile=True)
    f.compute_kernel(Xtr, Xte); st = f.stats()
print("exact dense:", json.dumps({k: round(st[k], 3) for k in st if k.startswith("ms_")}))
PY
