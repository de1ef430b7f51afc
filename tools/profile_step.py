"""Profiling driver (run under ncu): the C4 workload, one warm-up batch and one profiled batch.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --batch 8
"""
import argparse
import time
import os
import sys
from math import comb

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastsk_b200 import FastSK, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--len", type=int, default=200)
ap.add_argument("--g", type=int, default=16)
ap.add_argument("--m", type=int, default=8)
ap.add_argument("--alphabet", type=int, default=4)
ap.add_argument("--acc-path", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--batches-per-call", type=int, default=1)
ap.add_argument("--wave", type=int, default=32)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--rows-threads", type=int, default=0)
ap.add_argument("--pad", type=int, default=0)
ap.add_argument("--seg-fused", type=int, default=0)
ap.add_argument("--heavy-tau", type=int, default=0)
ap.add_argument("--heavy-cap", type=int, default=0)
ap.add_argument("--gemm-shape", type=int, default=0)
ap.add_argument("--skew", type=int, default=0, help="1 = first-order Markov GC-rich DNA with a planted 12-mer in half of the sequences (SURVEY 8d)")
ap.add_argument("--acc-unroll", type=int, default=2)
ap.add_argument("--acc-prefetch", type=int, default=1)
a = ap.parse_args()

X = np.random.default_rng(0).integers(1, a.alphabet + 1, size=(a.n, a.len), dtype=np.int32)
if a.skew:
    rng = np.random.default_rng(1)
    P = np.array([[0.10, 0.40, 0.40, 0.10], [0.05, 0.45, 0.45, 0.05], [0.05, 0.45, 0.45, 0.05], [0.10, 0.40, 0.40, 0.10]])
    cdf = np.cumsum(P, axis=1)
    X = np.empty((a.n, a.len), dtype=np.int32)
    X[:, 0] = rng.integers(0, 4, size=a.n)
    u = rng.random((a.n, a.len))
    for t in range(1, a.len):
        X[:, t] = (u[:, t, None] > cdf[X[:, t - 1]]).sum(axis=1)
    motif = rng.integers(0, 4, size=12)
    for i in np.flatnonzero(rng.random(a.n) < 0.5):
        p = int(rng.integers(0, a.len - 12 + 1))
        X[i, p:p + 12] = motif
    X = (np.minimum(X, 3) + 1).astype(np.int32)
order = np.random.default_rng(0).permutation(comb(a.g, a.m)).astype(np.int32)
f = FastSK(a.g, a.m, combo_sequence=order, distributed=False, profile=True)
f.set_option("batch", a.batch)
f.set_option("acc_path", a.acc_path)
f.set_option("wave", a.wave)
f.set_option("overlap", a.overlap)
f.set_option("rows_threads", a.rows_threads)
f.set_option("pad", a.pad)
f.set_option("seg_fused", a.seg_fused)
f.set_option("heavy_tau", a.heavy_tau)
f.set_option("heavy_cap", a.heavy_cap)
f.set_option("gemm_shape", a.gemm_shape)
f.set_option("acc_unroll", a.acc_unroll)
f.set_option("acc_prefetch", a.acc_prefetch)
codes = np.ascontiguousarray(X.reshape(-1))
offsets = np.arange(a.n + 1, dtype=np.int64) * a.len
f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), int(a.n * 0.8), a.n - int(a.n * 0.8))
for r in range(1 + a.reps):
    per = a.batch * a.batches_per_call
    q = np.ascontiguousarray(np.resize(order, (1 + a.reps) * per)[r * per:(r + 1) * per])
    f._call("fsk_accumulate_combos", q.ctypes.data_as(_lib.c_i32p), len(q), 1)
    if r == 0:
        s0 = f.stats()
        t0 = time.perf_counter()
wall_ms = (time.perf_counter() - t0) * 1e3
s1 = f.stats()
d = {k: s1[k] - s0[k] for k in s1 if k.startswith("ms_") or k in ("pair_updates", "entries", "runs", "combos_done", "kernel_launches", "heavy_runs")}
d["heavy_tau"] = s1["heavy_tau"]
d["wall_ms"] = wall_ms
d["combos_per_s"] = d["combos_done"] / wall_ms * 1e3
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
print({k: s1[k] for k in ("nfeat", "n_pairs", "key_bits", "id_bits", "record_bytes", "sort_passes", "batch")})
