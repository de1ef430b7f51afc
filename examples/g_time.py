#!/usr/bin/env python
"""Kernel time against g -- the reference's `results/run_experiments.py --g-time` sweep (g_time_experiment, :326-470) on this
backend: k = g - m = 6 fixed, g = 6 ... 20, and for every g the three FastSK configurations the reference times,
    FastSK-Exact                                  FastSK(g, m, t=20)
    FastSK-Approx 1 thread                        FastSK(g, m, t=1, approx=True, max_iters=C(g, m))   (runs until convergence)
    FastSK-Approx 20 thread no variance 50 iters  FastSK(g, m, t=20, approx=True, max_iters=50, skip_variance=True)
each on `compute_train` of the dataset's train file, as the reference's time_fastsk does.  Writes the reference's CSV columns.

    python examples/g_time.py --dataset EP300 [--min-g 6 --max-g 20 --output-dir .]
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


def time_fastsk(Xtrain, g, m, **kw):
    t0 = time.perf_counter()
    f = FastSK(g=g, m=m, **kw)
    f.compute_train(Xtrain)
    f.get_train_kernel()
    dt = time.perf_counter() - t0
    st = f.stats()
    return dt, st["combos_done"], st["acc_path"]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="EP300")
    ap.add_argument("--min-g", type=int, default=6)
    ap.add_argument("--max-g", type=int, default=20)
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--output-dir", default=".")
    a = ap.parse_args()
    Xtrain, _ = FastaUtility().read_data(os.path.join(ROOT, "data", a.dataset + ".train.fasta"))
    shortest = min(len(x) for x in Xtrain)
    time_fastsk(Xtrain, a.k + 1, 1)                       # context, allocator and module load are not part of the first row
    rows = []
    for g in range(max(a.min_g, a.k + 1), min(a.max_g, shortest) + 1):   # m = 0 (g = k) has no mismatch position to drop
        m = g - a.k
        exact, n_exact, path = time_fastsk(Xtrain, g, m, t=20)
        approx1, n1, _ = time_fastsk(Xtrain, g, m, t=1, approx=True, max_iters=comb(g, m))
        approx20, n20, _ = time_fastsk(Xtrain, g, m, t=20, approx=True, max_iters=50, skip_variance=True)
        rows.append({"g": g, "k": a.k, "m": m, "FastSK-Exact": exact, "FastSK-Approx 1 thread": approx1,
                     "FastSK-Approx 20 thread no variance 50 iters": approx20, "combinations": comb(g, m),
                     "combinations_exact": n_exact, "iterations_approx_1_thread": n1, "combinations_approx_20_thread": n20,
                     "acc_path": path})
        print(rows[-1], flush=True)
    out = os.path.join(a.output_dir, a.dataset + "_g_times.csv")
    with open(out, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0]))
        w.writeheader()
        w.writerows(rows)
    print("wrote", out)
