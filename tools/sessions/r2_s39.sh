#!/bin/bash
# r2 session 39 (1 GPU): byte operands in the exact dense contraction too (syrk_tc_kernel<NA, true>) -- dense tests, EP300 exact timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense or approx or variance or speculated or registers or golden or fingerprints" > gpurun_out/r2s39_pytest.txt 2>&1
tail -4 gpurun_out/r2s39_pytest.txt
timeout 300 python - > gpurun_out/r2s39_ep300_exact.txt 2>&1 <<'PY'
import json, sys, time
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for u8 in (1, 0):
    for shape in (0, 1, 2):
        best = None
        for rep in range(4):
            f = FastSK(10, 6, seed=0, device=0, distributed=False, profile=True)
            f.set_option("dense_u8", u8); f.set_option("gemm_shape", shape)
            f.compute_kernel(Xtr, Xte); st = f.stats()
            n, kdim = st["n_seq"], 256 * st["combos_done"]
            row = {"dense_u8": u8, "gemm_shape": shape, "device_ms": round(st["ms_total"], 3), "ms_accumulate": round(st["ms_accumulate"], 3), "ms_pack": round(st["ms_pack"], 3),
                   "tensor_tops": round(2.0 * (n * (n + 128) / 2.0) * kdim / (st["ms_accumulate"] * 1e-3) / 1e12, 1), "launches": st["kernel_launches"]}
            if best is None or row["device_ms"] < best["device_ms"]: best = row
        print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s39_ep300_exact.txt
