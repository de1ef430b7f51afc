#!/bin/bash
# r2 session 10 (1 GPU): GPU SVM test, example scripts, ncu of the speculated Welford contraction
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linear_svm or errors_and_api" > gpurun_out/r2s10_pytest.txt 2>&1
tail -4 gpurun_out/r2s10_pytest.txt
timeout 300 python examples/run.py --trn data/EP300.train.fasta --tst data/EP300.test.fasta -g 10 -m 6 -a -I 50 -t 1 --seed 0 --gpu-svm > gpurun_out/r2s10_run_gpu.txt 2>&1
tail -3 gpurun_out/r2s10_run_gpu.txt
timeout 300 python examples/run.py --trn data/EP300.train.fasta --tst data/EP300.test.fasta -g 10 -m 6 --seed 0 > gpurun_out/r2s10_run_cpu.txt 2>&1
tail -3 gpurun_out/r2s10_run_cpu.txt
timeout 900 python examples/stdev_iters.py --dataset EP300 -g 10 -m 4 --max-I 50 --gpu-svm --compare --output-dir gpurun_out > gpurun_out/r2s10_stdev_iters.txt 2>&1
cat gpurun_out/r2s10_stdev_iters.txt | tail -16
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_tc_welford -c 2 -f -o gpurun_out/r2s10_welford python /tmp/ep300_approx.py > gpurun_out/r2s10_ncu.log 2>&1
tail -2 gpurun_out/r2s10_ncu.log
