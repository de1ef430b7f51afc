#!/usr/bin/env python
"""Diagnostic for the heavy-run tensor-core stage: heavy_tau on against off, with a mismatch report."""
import os
import sys
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsk_b200 import FastSK  # noqa: E402


def unpack(K, n):
    M = np.zeros((n, n), dtype=np.int64)
    M[np.tril_indices(n)] = K
    return M


def run(n, L, g, m, nq, tau, batch, alpha=4, lowc=True, cap=0):
    rng = np.random.default_rng(0)
    X = []
    for _ in range(n):
        if lowc and rng.random() < 0.5:
            X.append(np.full(L, int(rng.integers(1, alpha + 1))).tolist())
        else:
            X.append(rng.integers(1, alpha + 1, size=L).tolist())
    queue = rng.permutation(comb(g, m))[:nq].astype(np.int32)
    out = {}
    for t in (-1, tau):
        f = FastSK(g, m, combo_sequence=queue, profile=True)
        f.set_option("acc_path", 2)
        f.set_option("heavy_tau", t)
        f.set_option("heavy_cap", cap)
        f.set_option("batch", batch)
        f.compute_train(X)
        out[t] = unpack(f.get_unnormalised(), n)
        st = f.stats()
        print(f"  tau={t}: heavy_runs={st['heavy_runs']} runs={st['runs']} entries={st['entries']} launches={st['kernel_launches']}", flush=True)
    a, b = out[-1], out[tau]
    bad = np.argwhere(a != b)
    print(f"n={n} L={L} g={g} m={m} nq={nq} tau={tau} batch={batch}: {len(bad)} of {n * (n + 1) // 2} cells differ; sum off {a.sum()} on {b.sum()}")
    for i, j in bad[:10]:
        print(f"    K[{i}][{j}]: off {a[i, j]} on {b[i, j]}")
    return len(bad)


if __name__ == "__main__":
    bad = 0
    bad += run(40, 30, 8, 4, 1, 8, 1)
    bad += run(40, 30, 8, 4, 4, 8, 4)
    bad += run(300, 60, 8, 4, 3, 8, 1)
    bad += run(300, 60, 16, 8, 3, 8, 1)
    bad += run(700, 100, 16, 8, 4, 8, 4, lowc=False)
    bad += run(700, 100, 16, 8, 4, 8, 4)
    sys.exit(1 if bad else 0)
