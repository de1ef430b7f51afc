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


def run(tag, name, g, m, reps=3, acc_path=0, **kw):
    """Best of `reps` runs; host phases by wall clock, device phases from the library's CUDA-event spans."""
    from fastsk_b200 import _lib
    from fastsk_b200.fastsk import _flatten
    Xtr, Ytr, Xte, Yte = load(name)
    best = None
    for r in range(reps):
        t0 = time.perf_counter()
        f = FastSK(g, m, seed=0, profile=True, **kw)
        f.set_option("acc_path", acc_path)
        ctr, otr = _flatten(Xtr)
        cte, ote = _flatten(Xte)
        codes = np.concatenate([ctr, cte])
        offsets = np.concatenate([otr, ote[1:] + otr[-1]])
        t1 = time.perf_counter()
        f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), len(otr) - 1, len(ote) - 1)
        t2 = time.perf_counter()
        f._call("fsk_build_partial")
        t3 = time.perf_counter()
        f._call("fsk_finalize")
        t4 = time.perf_counter()
        Ktr, Kte = f.get_train_kernel(), f.get_test_kernel()
        t5 = time.perf_counter()
        st = f.stats()
        row = {"config": tag, "data": name, "g": g, "m": m, **{k: v for k, v in kw.items()}, "n_seq": st["n_seq"], "nfeat": st["nfeat"],
               "combinations_total": comb(g, m), "combinations_done": st["combos_done"], "acc_path": st["acc_path"],
               "host_s": {"lists_to_flat": round(t1 - t0, 5), "upload": round(t2 - t1, 5), "build_partial": round(t3 - t2, 5),
                          "finalize": round(t4 - t3, 5), "getters_d2h": round(t5 - t4, 5), "total": round(t5 - t0, 5)},
               "device_ms": {k[3:]: round(st[k], 4) for k in st if k.startswith("ms_")},
               "combinations_per_s_build": st["combos_done"] / (t3 - t2),
               "combinations_per_s_device": st["combos_done"] / (st["ms_total"] * 1e-3) if st["ms_total"] else None,
               "stdevs": len(f.get_stdevs()), "kernel_launches": st["kernel_launches"], "batch": st["batch"],
               "record_bytes": st["record_bytes"], "sort_passes": st["sort_passes"]}
        if best is None or row["host_s"]["total"] < best["host_s"]["total"]:
            best = row
        del f
    print(json.dumps(best), flush=True)


if __name__ == "__main__":
    for path, ptag in ((2, "rows"), (3, "dense tensor-core")):
        run(f"C2 exact [{ptag}]", "EP300", 10, 6, t=20, acc_path=path)
        run(f"C2 approx t=1 max_iters=50 [{ptag}]", "EP300", 10, 6, acc_path=path, t=1, approx=True, max_iters=50)
        run(f"C2 approx t=20 max_iters=10 [{ptag}]", "EP300", 10, 6, acc_path=path, t=20, approx=True, max_iters=10)
    run("C2 approx t=20 skip_variance max_iters=50 [auto]", "EP300", 10, 6, t=20, approx=True, max_iters=50, skip_variance=True)
    run("C3 exact [auto]", "1.1", 10, 6, t=20)
    run("C5 approx t=1 max_iters=100 [auto]", "AImed", 20, 10, t=1, approx=True, max_iters=100)
    run("C5 approx t=20 max_iters=50 [auto]", "AImed", 20, 10, t=20, approx=True, max_iters=50)
