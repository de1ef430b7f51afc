#!/usr/bin/env python
"""Profiling driver for the variance mode (run under ncu): EP300-shaped synthetic input, t streams, a few iterations."""
import argparse
import os
import sys
from math import comb

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastsk_b200 import FastSK  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4000)
ap.add_argument("--len", type=int, default=100)
ap.add_argument("--g", type=int, default=10)
ap.add_argument("--m", type=int, default=6)
ap.add_argument("--t", type=int, default=20)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--acc-path", type=int, default=0)
a = ap.parse_args()
X = np.random.default_rng(0).integers(1, 5, size=(a.n, a.len), dtype=np.int32)
f = FastSK(a.g, a.m, a.t, True, 0.025, a.iters, False, seed=0, profile=True)
f.set_option("acc_path", a.acc_path)
f.compute_train(X)
st = f.stats()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k.startswith("ms_") or k in ("combos_done", "kernel_launches", "acc_path", "batch")})
