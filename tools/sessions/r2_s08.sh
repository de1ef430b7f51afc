#!/bin/bash
# r2 session 8 (1 GPU): variance mode with speculated rounds -- whole GPU suite, EP300 t=1 approx timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s08_pytest.txt 2>&1
tail -6 gpurun_out/r2s08_pytest.txt
timeout 300 python - > gpurun_out/r2s08_ep300_approx.txt 2>&1 <<'PY'
import json, sys, time
sys.path.insert(0, ".")
import bench
for depth in (1, 0):
    import fastsk_b200
    from fastsk_b200 import FastSK, FastaUtility
    fu = FastaUtility()
    Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
    for path in (3, 2):
        best = None
        for rep in range(3):
            f = FastSK(10, 6, t=1, approx=True, max_iters=50, seed=0, device=0, distributed=False, profile=True)
            f.set_option("acc_path", path); f.set_option("spec_depth", depth)
            t0 = time.perf_counter(); f.compute_kernel(Xtr, Xte); wall = time.perf_counter() - t0
            st = f.stats()
            row = {"spec_depth": depth, "acc_path": path, "wall_ms": round(wall * 1e3, 2), "device_ms": round(st["ms_total"], 3), "combos": st["combos_done"],
                   "combos_per_s_device": round(st["combos_done"] / (st["ms_total"] * 1e-3)), "launches": st["kernel_launches"], "stdevs": len(f.get_stdevs()),
                   "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")}}
            if best is None or row["device_ms"] < best["device_ms"]: best = row
        print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s08_ep300_approx.txt
