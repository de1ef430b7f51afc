#!/bin/bash
# r2 session 50 (1 GPU): the Welford contraction at a size whose running means do not fit L2 (16 000 sequences: 1 GB of means)
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2s50_welford_16k.txt 2>&1 <<'PY'
import json, sys
import numpy as np
sys.path.insert(0, ".")
from fastsk_b200 import FastSK
rng = np.random.default_rng(0)
for n in (8000, 16000):
    X = rng.integers(1, 5, size=(n, 100), dtype=np.int32)
    for regs, u8 in ((1, 1), (1, 0), (0, 0)):
        best = None
        for rep in range(2):
            f = FastSK(10, 6, t=1, approx=True, max_iters=20, delta=1e-9, seed=0, device=0, distributed=False, profile=True)
            f.set_option("acc_path", 3); f.set_option("wf_regs", regs); f.set_option("dense_u8", u8)
            f.compute_train(X); st = f.stats()
            row = {"n": n, "wf_regs": regs, "dense_u8": u8, "combos": st["combos_done"], "ms_accumulate": round(st["ms_accumulate"], 3), "ms_total": round(st["ms_total"], 3), "dense_mode": st["dense_mode"]}
            if best is None or row["ms_total"] < best["ms_total"]: best = row
            del f
        print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s50_welford_16k.txt
