#!/bin/bash
# r2 session 48 (1 GPU): ncu --set full of the exact dense contraction with byte operands (EP300, all 210 combinations in one launch)
mkdir -p gpurun_out
cat > /tmp/ep300_exact.py <<'PY'
import sys
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for rep in range(2):
    f = FastSK(10, 6, seed=0, device=0, distributed=False)
    f.compute_kernel(Xtr, Xte)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"syrk_tc_kernel|dense_count" -c 4 -f -o gpurun_out/r2s48_dense_u8 python /tmp/ep300_exact.py > gpurun_out/r2s48_ncu.log 2>&1
tail -2 gpurun_out/r2s48_ncu.log
