#!/bin/bash
# r2 session 40 (1 GPU): one or two output tiles per CTA in the exact dense contraction, by N and operand type
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/r2s40_gemm_shape.txt 2>&1 <<'PY'
import json, sys
import numpy as np
sys.path.insert(0, ".")
from fastsk_b200 import FastSK
rng = np.random.default_rng(0)
for n in (1000, 2000, 4000, 8000, 16000, 32000):
    X = rng.integers(1, 5, size=(n, 100), dtype=np.int32)
    for u8 in (1, 0):
        row = {"n": n, "dense_u8": u8}
        for shape in (1, 2):
            best = None
            for rep in range(3):
                f = FastSK(10, 6, seed=0, device=0, distributed=False, profile=True)
                f.set_option("acc_path", 3); f.set_option("dense_u8", u8); f.set_option("gemm_shape", shape)
                f.compute_train(X); st = f.stats()
                best = st["ms_accumulate"] if best is None else min(best, st["ms_accumulate"])
                del f
            row[f"shape{shape}_ms"] = round(best, 3)
        row["tops_best"] = round(2.0 * (n * (n + 128) / 2.0) * 256 * 210 / (min(row["shape1_ms"], row["shape2_ms"]) * 1e-3) / 1e12, 1)
        print(json.dumps(row), flush=True)
PY
cat gpurun_out/r2s40_gemm_shape.txt
