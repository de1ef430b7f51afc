#!/bin/bash
# r2 session 16 (1 GPU): ncu of the Welford contraction after the epilogue changes
mkdir -p gpurun_out
cat > /tmp/ep300_approx.py <<'PY'
import sys
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for rep in range(2):
    f = FastSK(10, 6, t=1, approx=True, max_iters=50, seed=0, device=0, distributed=False)
    f.compute_kernel(Xtr, Xte)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_tc_welford -c 1 -f -o gpurun_out/r2s16_welford python /tmp/ep300_approx.py > gpurun_out/r2s16_ncu.log 2>&1
tail -2 gpurun_out/r2s16_ncu.log
