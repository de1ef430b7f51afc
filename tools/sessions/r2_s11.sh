#!/bin/bash
# r2 session 11 (1 GPU): Welford with the reciprocal division (self-test, variance tests, EP300 t=1 timing), accumulate experiments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "division or approx or variance or speculated or linear_svm or golden or column_windows" > gpurun_out/r2s11_pytest.txt 2>&1
tail -4 gpurun_out/r2s11_pytest.txt
timeout 300 python - > gpurun_out/r2s11_ep300_approx.txt 2>&1 <<'PY'
import json, sys, time
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for path in (3, 2):
    best = None
    for rep in range(3):
        f = FastSK(10, 6, t=1, approx=True, max_iters=50, seed=0, device=0, distributed=False, profile=True)
        f.set_option("acc_path", path)
        t0 = time.perf_counter(); f.compute_kernel(Xtr, Xte); wall = time.perf_counter() - t0
        st = f.stats()
        row = {"acc_path": path, "wall_ms": round(wall * 1e3, 2), "device_ms": round(st["ms_total"], 3), "combos": st["combos_done"],
               "combos_per_s_device": round(st["combos_done"] / (st["ms_total"] * 1e-3)), "launches": st["kernel_launches"],
               "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")}}
        if best is None or row["device_ms"] < best["device_ms"]: best = row
    print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s11_ep300_approx.txt
timeout 900 python tools/c4_steps.py '{"count_updates": 0, "fit_smem": 0}' '{"count_updates": 0, "fit_smem": 1}' '{"count_updates": 0, "pf_stride": 64}' '{"count_updates": 0, "fit_smem": 1, "rows_threads": 512}' > gpurun_out/r2s11_steps.txt 2>&1
cat gpurun_out/r2s11_steps.txt
