#!/usr/bin/env python
"""Kernel time against the degree of parallelism -- the reference's thread experiment (results/run_experiments.py:114-160:
for t = 1 ... 20 time the train kernel of one dataset; only the `fastsk_I50` column -- approx, t=1, max_iters=50 -- is
live there).  The reference's t is host threads; here the resource that scales is GPUs, and `t` is only the number of
virtual streams of the approximation, so the sweep is over the GPUs one plain process drives (fsk_set_devices):

    fastsk_exact_time      FastSK(g, m)                                        exact, every combination
    fastsk_approx_time     FastSK(g, m, t=20, approx=True, max_iters=50)       20 streams with the variance test
    fastsk_approx_time_t1  FastSK(g, m, t=1, approx=True, max_iters=C(g, m))   one stream to convergence
    fastsk_I50             FastSK(g, m, t=1, approx=True, max_iters=50)        the reference's live column

    python examples/thread_time.py --dataset EP300 -g 10 -m 6
"""
import argparse
import csv
import os
import sys
import time
from math import comb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsk_b200 import FastSK, FastaUtility  # noqa: E402


def time_fastsk(X, g, m, devices, **kw):
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        f = FastSK(g=g, m=m, devices=devices, distributed=False, seed=0, **kw)
        f.compute_train(X)
        f.get_train_kernel()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        del f
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="EP300")
    ap.add_argument("-g", type=int, default=10)
    ap.add_argument("-m", type=int, default=6)
    ap.add_argument("--output-dir", default=".")
    a = ap.parse_args()
    import ctypes
    n_gpus = ctypes.c_int(0)
    try:
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n_gpus))
    except OSError:
        import torch
        n_gpus.value = torch.cuda.device_count()
    X, _ = FastaUtility().read_data(os.path.join(ROOT, "data", a.dataset + ".train.fasta"))
    rows = []
    for n in range(1, max(1, n_gpus.value) + 1):
        dev = list(range(n))
        row = {"gpus": n,
               "fastsk_exact_time": time_fastsk(X, a.g, a.m, dev),
               "fastsk_approx_time": time_fastsk(X, a.g, a.m, dev, t=20, approx=True, max_iters=50),
               "fastsk_approx_time_t1": time_fastsk(X, a.g, a.m, dev, t=1, approx=True, max_iters=comb(a.g, a.m)),
               "fastsk_I50": time_fastsk(X, a.g, a.m, dev, t=1, approx=True, max_iters=50)}
        rows.append(row)
        print(a.dataset, row, flush=True)
    out = os.path.join(a.output_dir, a.dataset + "_vary_gpus_I50.csv")
    with open(out, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0]))
        w.writeheader()
        w.writerows(rows)
    print("wrote", out)
