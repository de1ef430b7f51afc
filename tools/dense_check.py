#!/usr/bin/env python
"""Diagnostic for the dense tensor-core path: acc_path 3 against acc_path 2 on random DNA, with a mismatch report."""
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


def main(n=300, L=80, g=10, m=6, nq=5):
    rng = np.random.default_rng(0)
    X = rng.integers(1, 5, size=(n, L), dtype=np.int32)
    queue = rng.permutation(comb(g, m))[:nq].astype(np.int32)
    out = {}
    for path in (2, 3):
        f = FastSK(g, m, combo_sequence=queue, profile=True)
        f.set_option("acc_path", path)
        f.compute_train(X)
        out[path] = unpack(f.get_unnormalised(), n)
        print(path, {k: v for k, v in f.stats().items() if k in ("acc_path", "batch", "kernel_launches", "ms_pack", "ms_accumulate")}, flush=True)
    a, b = out[2], out[3]
    bad = np.argwhere(a != b)
    print(f"n={n} nq={nq}: {len(bad)} of {n * (n + 1) // 2} cells differ")
    if len(bad):
        for i, j in bad[:12]:
            print(f"  K[{i}][{j}]: rows {a[i, j]}  dense {b[i, j]}")
        ti, tj = bad[:, 0] // 128, bad[:, 1] // 128
        print("  tiles with mismatches:", sorted(set(zip(ti.tolist(), tj.tolist()))))
        print("  sum rows", a.sum(), "sum dense", b.sum())
        sys.exit(1)


if __name__ == "__main__":
    main()
    main(n=100, nq=1)
    main(n=520, nq=96)
