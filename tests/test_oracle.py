"""The oracle against the reference's golden vectors (CPU only).

Pins oracle/fsk_oracle.c (the C restatement) to
  * fixtures generated from the unmodified reference engine (tests/golden/*.npz, make_golden.py),
  * the known-answer vectors of SURVEY.md section 8c,
  * oracle/_ref itself, when it was built in this checkout,
  * the Hamming-distance identity (an independent NumPy statement).
"""
from math import comb

import numpy as np
import pytest

from conftest import golden_names, load_golden


def run_c(oracle, d, normalise):
    return oracle.run("c", d["Xtrain"], d["Xtest"], d["g"], d["m"], d["queue"], T=d["T"], approx=d["approx"],
                      delta=d["delta"], max_iters=d["max_iters"], skip_variance=d["skip_variance"], normalise=normalise)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_reference_fixture(oracle_mod, name):
    d = load_golden(name)
    K, Ki, sd = run_c(oracle_mod, d, False)
    Kn, _, sd2 = run_c(oracle_mod, d, True)
    welford = d["approx"] and not d["skip_variance"]
    if welford and d["T"] > 2:
        # the reference adds its T per-thread means in lock-acquisition order (fastsk_kernel.cpp:296-309):
        # an fp64 sum of >2 terms is order dependent in the last bit
        np.testing.assert_allclose(K, d["K_un"], rtol=1e-14, atol=0)
        np.testing.assert_allclose(Kn, d["K_norm"], rtol=1e-14, atol=0)
    else:
        assert np.array_equal(K, d["K_un"])
        assert np.array_equal(Kn, d["K_norm"])
    if not welford:
        assert np.array_equal(Ki.astype(np.float64), d["K_un"])
    assert sd == d["stdevs"].tolist() and sd2 == sd
    if welford:
        assert sd[0] == 3162.2775020544923        # sqrt(9999999), fastsk_kernel.cpp:134-136,247


def test_kat_small_g3_m1(oracle_mod):
    # data/small.*  ACACA,AAACA | ACACA,AACCA with first-seen ids
    tr, te = [[1, 2, 1, 2, 1], [1, 1, 1, 2, 1]], [[1, 2, 1, 2, 1], [1, 1, 2, 2, 1]]
    K, Ki, _ = oracle_mod.run("c", tr, te, 3, 1, [0, 1, 2])
    S = oracle_mod.unpack(Ki, 4)
    assert S[:2, :2].tolist() == [[15, 9], [9, 13]]
    Kn, _, _ = oracle_mod.run("c", tr, te, 3, 1, [2, 0, 1], normalise=True)
    Sn = oracle_mod.unpack(Kn, 4)
    assert Sn[1, 0] == 0.6445033866354896
    assert Sn[2:, :2].tolist() == [[1.0, 0.6445033866354896], [0.3892494720807615, 0.5853694070049635]]


def test_kat_docs_toy_g3_m2(oracle_mod):
    # docs/2demo/fastDemo.ipynb cell 3
    tr, te = [[1, 0, 1, 0, 1], [1, 1, 1, 0, 1]], [[1, 1, 1, 1, 1], [1, 0, 1, 0, 1]]
    Kn, _, sd = oracle_mod.run("c", tr, te, 3, 2, [0, 1, 2], normalise=True)
    Sn = oracle_mod.unpack(Kn, 4)
    assert Sn[1, 0] == 0.8885233166386385
    assert Sn[2:, :2].tolist() == [[0.7453559924999299, 0.9271726499455306], [1.0, 0.8885233166386385]]
    assert sd == []


def test_kat_dataset_heads(oracle_mod):
    # first three train sequences of the bundled datasets (K_ij depends only on seqs i, j)
    import os
    from conftest import DATA_DIR
    from fastsk_b200.utils import FastaUtility
    want = {
        ("EP300", 10, 6): ([[27006, 6869, 8715], [6869, 26090, 5594], [8715, 5594, 25988]],
                           (0.25877740004582855, 0.3289658610827968, 0.21483201081638156)),
        ("1.1", 10, 6): ([[28650, 3359, 4967], [3359, 29424, 6665], [4967, 6665, 29628]],
                         (0.11569027002773254, 0.17048284289040663, 0.22573459902716578)),
        ("AImed", 8, 4): ([[12436, 11890, 11474], [11890, 12436, 11535], [11474, 11535, 13136]],
                          (0.9560952074622064, 0.8977241717470419, 0.9024968033033056)),
    }
    for (name, g, m), (K_want, norm_want) in want.items():
        X, _ = FastaUtility().read_data(os.path.join(DATA_DIR, name + ".train.fasta"))
        lut = {}
        X3 = [[lut.setdefault(v, len(lut) + 1) for v in x] for x in X[:3]]
        q = list(range(comb(g, m)))
        K, Ki, _ = oracle_mod.run("c", X3, [], g, m, q)
        assert oracle_mod.unpack(Ki, 3).tolist() == K_want
        Kn, _, _ = oracle_mod.run("c", X3, [], g, m, q, normalise=True)
        assert (Kn[1], Kn[3], Kn[4]) == norm_want


def test_combination_order_is_lexicographic(oracle_mod):
    from itertools import combinations
    for g, k in [(3, 2), (5, 3), (10, 4), (8, 8), (7, 1)]:
        want = [list(c) for c in combinations(range(g), k)]   # shared.cpp:347-360 emits this order
        got = [oracle_mod.combination(g, k, i) for i in range(comb(g, k))]
        assert got == want
        assert oracle_mod.c_lib().fsko_nchoosek(g, k) == comb(g, k)


@pytest.mark.parametrize("seed,alpha,g,m", [(0, 4, 6, 2), (1, 2, 5, 3), (2, 20, 4, 1), (3, 4, 8, 4)])
def test_hamming_identity(oracle_mod, seed, alpha, g, m):
    rng = np.random.default_rng(seed)
    X = [rng.integers(1, alpha + 1, size=int(rng.integers(g, g + 12))).tolist() for _ in range(9)]
    _, Ki, _ = oracle_mod.run("c", X[:6], X[6:], g, m, rng.permutation(comb(g, m)))
    assert np.array_equal(oracle_mod.unpack(Ki, 9).astype(np.int64), oracle_mod.hamming_kernel(X, g, m))


def test_partial_kernels_sum_to_exact(oracle_mod):
    rng = np.random.default_rng(7)
    g, m = 6, 3
    X = [rng.integers(1, 5, size=int(rng.integers(g, 30))).tolist() for _ in range(12)]
    total = np.zeros(12 * 13 // 2, dtype=np.uint64)
    for c in range(comb(g, m)):
        total += oracle_mod.partial(X, g, oracle_mod.combination(g, g - m, c))
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, list(range(comb(g, m))))
    assert np.array_equal(total, Ki)


def test_c_oracle_matches_live_reference(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built in this checkout")
    rng = np.random.default_rng(11)
    for alpha, g, m, T, approx, mi, skip in [(4, 8, 4, 3, False, -1, False), (21, 6, 2, 2, True, 6, False),
                                              (5, 7, 3, 1, True, -1, False), (4, 9, 5, 4, True, 5, True)]:
        X = [rng.integers(1, alpha + 1, size=int(rng.integers(g, 60))).tolist() for _ in range(25)]
        lut = {}
        X = [[lut.setdefault(v, len(lut) + 1) for v in x] for x in X]
        q = rng.permutation(comb(g, m)).astype(np.int32)
        a = oracle_mod.run("c", X[:15], X[15:], g, m, q, T=T, approx=approx, max_iters=mi, skip_variance=skip, normalise=True)
        b = oracle_mod.run("ref", X[:15], X[15:], g, m, q, T=T, approx=approx, max_iters=mi, skip_variance=skip, normalise=True)
        if approx and not skip and T > 2:
            np.testing.assert_allclose(a[0], b[0], rtol=1e-14)
        else:
            assert np.array_equal(a[0], b[0])
        assert a[2] == b[2]


def test_oracle_rejects_short_sequences(oracle_mod):
    with pytest.raises(RuntimeError):
        oracle_mod.run("c", [[1, 2, 3]], [[1, 2]], 3, 1, [0, 1, 2])
