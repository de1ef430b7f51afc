"""Fingerprints of the FULL bundled configurations from the UNMODIFIED reference engine (oracle/_ref).

Run in the authoring container only:   python tests/golden/make_fingerprints.py
For each configuration of BASELINE.json that the reference can finish here (EP300 g=10 m=6 exact: all 4000 sequences,
210 combinations, ~80 s on 8 cores; protein 1.1 g=10 m=6 exact: 3574 sequences) it stores in tests/golden/fingerprints.json
the sha256 of the unnormalised kernel (int64 packed lower triangle) and of the normalised one (fp64), the 3 x 3 known-answer
block of SURVEY.md 8c and a few sampled cells -- a few hundred bytes that pin ~8 million cells each on the GPU
(tests/test_gpu_parity.py::test_full_bundled_sets_against_the_reference_fingerprints)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fastsk_b200.utils import FastaUtility  # noqa: E402

DATA = os.path.join(ROOT, "data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fingerprints.json")


def main():
    res = {}
    for name, g, m in (("1.1", 10, 6), ("EP300", 10, 6)):
        fu = FastaUtility()
        tr, _ = fu.read_data(os.path.join(DATA, name + ".train.fasta"))
        te, _ = fu.read_data(os.path.join(DATA, name + ".test.fasta"))
        n = len(tr) + len(te)
        from math import comb
        queue = np.arange(comb(g, m), dtype=np.int32)
        t0 = time.time()
        K_un, _, _ = oracle.run("ref", tr, te, g, m, queue, T=os.cpu_count() or 1, normalise=False)
        K_n = oracle.normalise(K_un, n)
        Ki = K_un.astype(np.int64)
        assert np.array_equal(Ki.astype(np.float64), K_un)
        rng = np.random.default_rng(7)
        cells = np.sort(rng.choice(Ki.size, size=16, replace=False))
        tri = lambda i, j: i * (i + 1) // 2 + j  # noqa: E731
        res[name] = {
            "g": g, "m": m, "n_train": len(tr), "n_test": len(te), "n_pairs": int(Ki.size), "reference_seconds": round(time.time() - t0, 1),
            "sha256_unnormalised_int64": hashlib.sha256(Ki.tobytes()).hexdigest(),
            "sha256_normalised_f64": hashlib.sha256(K_n.tobytes()).hexdigest(),
            "kat_3x3_unnormalised": [[int(Ki[tri(max(a, b), min(a, b))]) for b in range(3)] for a in range(3)],
            "kat_normalised_10_20_21": [float(K_n[tri(1, 0)]), float(K_n[tri(2, 0)]), float(K_n[tri(2, 1)])],
            "cells": [int(c) for c in cells], "cells_unnormalised": [int(Ki[c]) for c in cells],
            "cells_normalised_hex": [float(K_n[c]).hex() for c in cells],
            "sum_unnormalised": int(Ki.sum()),
        }
        print(name, res[name]["reference_seconds"], "s", res[name]["kat_3x3_unnormalised"], flush=True)
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
