"""Generate tests/golden/*.npz from the UNMODIFIED reference engine (oracle/_ref).

Run in the authoring container only (needs /root/reference to have been present when
oracle/_ref was built):   python tests/golden/make_golden.py
Each fixture stores the inputs (flat codes/offsets, parameters, the fixed combination queue)
and what the reference produced for them: unnormalised K (fp64 packed lower triangle, read
before fastsk_kernel.cpp:96-103), normalised K, and stdevs.
"""
import os
import sys
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fastsk_b200.utils import FastaUtility  # noqa: E402

DATA = os.path.join(ROOT, "data")
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, ntr, nte):
    r = FastaUtility()
    a, _ = r.read_data(os.path.join(DATA, name + ".train.fasta"))
    b, _ = r.read_data(os.path.join(DATA, name + ".test.fasta"))
    return a[:ntr], b[:nte]


def recode(tr, te):
    """dense first-seen ids from 1 (what a fresh FastaUtility would hand out for this subset)."""
    lut = {}
    out = []
    for X in (tr, te):
        out.append([[lut.setdefault(v, len(lut) + 1) for v in x] for x in X])
    return out


def emit(name, tr, te, g, m, T=1, approx=False, delta=0.025, max_iters=-1, skip_variance=False, queue=None, seed=0):
    nc = comb(g, m)
    if queue is None:
        queue = np.random.default_rng(seed).permutation(nc).astype(np.int32)
    queue = np.asarray(queue, dtype=np.int32)
    kw = dict(T=T, approx=approx, delta=delta, max_iters=max_iters, skip_variance=skip_variance)
    K_un, _, sd = oracle.run("ref", tr, te, g, m, queue, normalise=False, **kw)
    K_n, _, sd2 = oracle.run("ref", tr, te, g, m, queue, normalise=True, **kw)
    assert sd == sd2
    codes, offsets = oracle.flatten(list(tr) + list(te))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), codes=codes, offsets=offsets, n_train=len(tr), n_test=len(te),
                        g=g, m=m, T=T, approx=approx, delta=delta, max_iters=max_iters, skip_variance=skip_variance,
                        queue=queue, K_un=K_un, K_norm=K_n, stdevs=np.asarray(sd, dtype=np.float64))
    print(f"{name}: N={len(tr) + len(te)} combos={len(queue)} stdevs={len(sd)}")


def main():
    tr, te = load("small", 2, 2)
    emit("small_g3m1", tr, te, 3, 1)
    emit("small_g5m2", tr, te, 5, 2)
    # docs/2demo/fastDemo.ipynb cell 3 toy (raw ids incl. 0)
    emit("toy_g3m2", [[1, 0, 1, 0, 1], [1, 1, 1, 0, 1]], [[1, 1, 1, 1, 1], [1, 0, 1, 0, 1]], 3, 2)
    tr, te = load("EP300", 48, 24)
    emit("ep300_exact", tr, te, 10, 6, T=4)
    emit("ep300_approx_t1", tr, te, 10, 6, T=1, approx=True)                      # test/run_check.py:45 settings
    emit("ep300_approx_conv", tr, te, 10, 6, T=2, approx=True, delta=2.0)          # convergence stop fires
    emit("ep300_approx_t3_i20", tr, te, 10, 6, T=3, approx=True, max_iters=20)
    emit("ep300_skipvar_t4_i10", tr, te, 10, 6, T=4, approx=True, max_iters=10, skip_variance=True)
    tr, te = recode(*load("1.1", 40, 20))
    emit("protein_exact", tr, te, 10, 6, T=3)
    emit("protein_g7m2_exact", tr, te, 7, 2, T=2)
    emit("protein_approx_t2", tr, te, 10, 6, T=2, approx=True, max_iters=30)
    tr, te = recode(*load("AImed", 24, 12))
    emit("aimed_g8m4_exact", tr, te, 8, 4, T=2)
    q = np.random.default_rng(5).choice(comb(20, 10), size=240, replace=False)
    emit("aimed_g20m10_skipvar", tr, te, 20, 10, T=4, approx=True, max_iters=60, skip_variance=True, queue=q)
    emit("aimed_g20m10_approx", tr, te, 20, 10, T=3, approx=True, max_iters=40, queue=q)


if __name__ == "__main__":
    main()
