#!/usr/bin/env python
"""Time the bundled BASELINE.json configurations through the public API on one GPU (host buffers in, kernels out).

    python tools/run_configs.py            # C2 EP300 exact + approx, C3 protein 1.1 exact, C5 AImed approx
Prints one JSON line per configuration: wall seconds of compute_kernel + getters, combinations processed, combinations/s.
"""
import json
import os
import sys
import time
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsk_b200 import FastSK, FastaUtility  # noqa: E402


def load(name):
    fu = FastaUtility()
    Xtr, Ytr = fu.read_data(os.path.join(ROOT, "data", f"{name}.train.fasta"))
    Xte, Yte = fu.read_data(os.path.join(ROOT, "data", f"{name}.test.fasta"))
    return Xtr, Ytr, Xte, Yte


def run(tag, name, g, m, reps=3, **kw):
    Xtr, Ytr, Xte, Yte = load(name)
    best = None
    for r in range(reps):
        t0 = time.perf_counter()
        f = FastSK(g, m, seed=0, **kw)
        f.compute_kernel(Xtr, Xte)
        t1 = time.perf_counter()
        Ktr, Kte = f.get_train_kernel(), f.get_test_kernel()
        t2 = time.perf_counter()
        st = f.stats()
        row = {"config": tag, "data": name, "g": g, "m": m, **{k: v for k, v in kw.items()}, "n_seq": st["n_seq"], "nfeat": st["nfeat"],
               "combinations_total": comb(g, m), "combinations_done": st["combos_done"], "compute_kernel_s": t1 - t0,
               "getters_s": t2 - t1, "combinations_per_s": st["combos_done"] / (t1 - t0), "stdevs": len(f.get_stdevs()),
               "kernel_launches": st["kernel_launches"], "batch": st["batch"], "record_bytes": st["record_bytes"],
               "sort_passes": st["sort_passes"]}
        if best is None or row["compute_kernel_s"] < best["compute_kernel_s"]:
            best = row
        del f
    print(json.dumps(best), flush=True)


if __name__ == "__main__":
    run("C2 exact", "EP300", 10, 6, t=20)
    run("C2 approx t=1 max_iters=50", "EP300", 10, 6, t=1, approx=True, max_iters=50)
    run("C2 approx t=20 skip_variance max_iters=50", "EP300", 10, 6, t=20, approx=True, max_iters=50, skip_variance=True)
    run("C3 exact", "1.1", 10, 6, t=20)
    run("C5 approx t=1 max_iters=100", "AImed", 20, 10, t=1, approx=True, max_iters=100)
    run("C5 approx t=20 max_iters=50", "AImed", 20, 10, t=20, approx=True, max_iters=50)
