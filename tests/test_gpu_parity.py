"""Parity of the CUDA path against the oracle, through the C ABI (needs a B200: -m gpu).

Bit-exact for the integer partial kernels and the unnormalised kernel; the fp64 normalised kernel
is compared bit-exactly in the integer modes (same IEEE mul/sqrt/div on both sides) and to 1e-12
relative in variance mode (the block-wise fp64 reduction order differs from the reference's
sequential sum; tolerance from BASELINE.json's north_star).
"""
import ctypes
import os
import zlib
from math import comb

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def FastSK():
    from fastsk_b200 import FastSK as cls
    return cls


def gpu_run(FastSK, d, **extra):
    f = FastSK(d["g"], d["m"], d["T"], d["approx"], d["delta"], d["max_iters"], d["skip_variance"],
               combo_sequence=d["queue"], **extra)
    if d["n_test"]:
        f.compute_kernel(d["Xtrain"], d["Xtest"])
    else:
        f.compute_train(d["Xtrain"])
    return f


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixture(FastSK, oracle_mod, name):
    d = load_golden(name)
    f = gpu_run(FastSK, d)
    n = d["n_train"] + d["n_test"]
    welford = d["approx"] and not d["skip_variance"]
    K_un = f.get_unnormalised(np.float64)
    Kp = f.get_kernel_packed()
    if welford:
        np.testing.assert_allclose(K_un, d["K_un"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(Kp, d["K_norm"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(f.get_stdevs(), d["stdevs"], rtol=RTOL, atol=0)
        assert len(f.get_stdevs()) == len(d["stdevs"])        # same stopping iteration
    else:
        assert np.array_equal(f.get_unnormalised(np.int64).astype(np.float64), d["K_un"])
        assert np.array_equal(K_un, d["K_un"])
        assert np.array_equal(Kp, d["K_norm"])
        assert f.get_stdevs() == []
    S = oracle_mod.unpack(Kp, n)
    assert np.array_equal(f.get_train_kernel(), S[:d["n_train"], :d["n_train"]])
    assert np.array_equal(f.get_test_kernel(), S[d["n_train"]:, :d["n_train"]])


def random_seqs(rng, n, alpha, lo, hi, low_complexity=False):
    X = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        if low_complexity and rng.random() < 0.5:
            X.append(np.full(L, int(rng.integers(1, alpha + 1))).tolist())      # one repeated character
        else:
            X.append(rng.integers(1, alpha + 1, size=L).tolist())
    return X


CASES = [
    # name, n_train, n_test, alphabet, len range, g, m, batch, low complexity
    ("dna_r32", 40, 17, 4, (12, 90), 8, 4, 0, False),
    ("dna_len_eq_g", 9, 4, 4, (10, 10), 10, 6, 0, False),
    ("dna_one_seq", 1, 0, 4, (30, 30), 6, 2, 0, False),
    ("dna_lowcomplex", 30, 10, 4, (20, 120), 7, 3, 3, True),
    ("dna_multi_tile", 300, 100, 4, (100, 100), 10, 6, 1, False),        # 36k windows: 9 sort tiles, look-back
    ("dna5_r64", 50, 20, 5, (40, 80), 16, 4, 0, False),                  # 12 x 3 = 36 key bits -> 64-bit records
    ("binary", 30, 10, 2, (25, 60), 12, 5, 2, True),
    ("protein_r32", 60, 25, 21, (16, 200), 7, 3, 0, False),              # 4 x 5 = 20 bits + 7 id bits
    ("protein_r64", 60, 25, 23, (30, 150), 12, 4, 0, False),             # 8 x 5 = 40 bits
    ("text_kv", 40, 15, 57, (25, 300), 20, 10, 0, False),                # 10 x 6 = 60 key bits + ids -> key/value sort, 2-word g-mers
    ("text_kv_multi_tile", 150, 50, 57, (60, 160), 20, 10, 4, False),
    ("bytes_b8", 25, 10, 200, (20, 90), 9, 2, 0, False),                 # 8-bit characters, 56 key bits
    ("wide_alphabet_ids", 20, 8, 30, (18, 50), 6, 3, 0, False),
]


def dense_eligible(alpha, g, m):
    """acc_path 3 (tensor-core contraction) takes at most 12 key bits: (g - m) x ceil(log2(alphabet))."""
    return (g - m) * max(1, int(np.ceil(np.log2(alpha)))) <= 12


@pytest.mark.parametrize("acc_path", [0, 1, 2, 3], ids=["auto", "global_red", "rows", "dense_tc"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_random_exact_vs_oracle(FastSK, oracle_mod, case, acc_path):
    name, ntr, nte, alpha, (lo, hi), g, m, batch, lowc = case
    if acc_path == 3 and not dense_eligible(alpha, g, m):
        pytest.skip("key space too large for the dense path")
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    X = random_seqs(rng, ntr + nte, alpha, max(lo, g), hi, lowc)
    if name == "wide_alphabet_ids":
        X = [[v * 1000003 for v in x] for x in X]        # sparse, huge ids: the library re-codes densely
    nc = comb(g, m)
    queue = rng.permutation(nc)[:min(nc, 60)].astype(np.int32)
    f = FastSK(g, m, combo_sequence=queue)
    if batch:
        f.set_option("batch", batch)
    f.set_option("acc_path", acc_path)
    f.compute_kernel(X[:ntr], X[ntr:]) if nte else f.compute_train(X[:ntr])
    lut = {}
    Xd = [[lut.setdefault(v, len(lut) + 1) for v in x] for x in X]
    K, Ki, _ = oracle_mod.run("c", Xd[:ntr], Xd[ntr:], g, m, queue)
    assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
    assert np.array_equal(f.get_kernel_packed(), oracle_mod.normalise(K, ntr + nte))
    st = f.stats()
    assert st["combos_done"] == len(queue) and st["kernel_launches"] > 0
    if acc_path:
        assert st["acc_path"] == acc_path


@pytest.mark.parametrize("case", [CASES[0], CASES[-3], CASES[-4]], ids=lambda c: c[0])
def test_safe_ranking_path_gives_the_same_kernel(FastSK, case):
    """onesweep's optimistic ranking (one shared atomic per key, verified by segment_kernel) and the match-mask
    ranking it falls back to must produce the same integers."""
    name, ntr, nte, alpha, (lo, hi), g, m, batch, lowc = case
    rng = np.random.default_rng(7)
    X = random_seqs(rng, ntr + nte, alpha, max(lo, g), hi, lowc)
    queue = rng.permutation(comb(g, m))[:40].astype(np.int32)
    out = []
    for safe in (0, 1):
        f = FastSK(g, m, combo_sequence=queue)
        f.set_option("safe_rank", safe)
        f.compute_train(X)
        out.append(f.get_unnormalised())
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("alpha,g,m", [(4, 8, 4), (21, 6, 2), (57, 12, 6)])
def test_per_combination_counts(FastSK, oracle_mod, alpha, g, m):
    """Integer partial kernel of every single combination (fastsk_kernel.cpp:224-241 for one work item)."""
    from fastsk_b200 import _lib
    rng = np.random.default_rng(alpha * 100 + g)
    X = random_seqs(rng, 24, alpha, g, 70)
    f = FastSK(g, m, combo_sequence=[0])
    f.compute_train(X)
    n = len(X)
    out = np.empty(n * (n + 1) // 2, dtype=np.int64)
    for c in rng.permutation(comb(g, m))[:12]:
        f._call("fsk_reset_partial")
        arr = np.array([c], dtype=np.int32)
        f._call("fsk_accumulate_combos", arr.ctypes.data_as(_lib.c_i32p), 1, 1)
        f._call("fsk_get_unnormalised_i64", out.ctypes.data_as(_lib.c_i64p))
        want = oracle_mod.partial(X, g, oracle_mod.combination(g, g - m, int(c)))
        assert np.array_equal(out.astype(np.uint64), want), f"combination {c}"


@pytest.mark.parametrize("T,max_iters,skip,delta", [(1, 12, False, 0.025), (3, 8, False, 0.025), (5, 4, True, 0.025),
                                                    (2, -1, False, 3.0), (-1, 3, False, 0.025), (4, -1, True, 0.025)])
@pytest.mark.parametrize("acc_path", [2, 3], ids=["rows", "dense_tc"])
def test_approx_modes_vs_oracle(FastSK, oracle_mod, T, max_iters, skip, delta, acc_path):
    rng = np.random.default_rng(42 + (T % 7) + 3 * max_iters)
    g, m = 9, 5
    X = random_seqs(rng, 36, 4, 20, 70)
    queue = rng.permutation(comb(g, m)).astype(np.int32)
    f = FastSK(g, m, T, True, delta, max_iters, skip, combo_sequence=queue)
    f.set_option("acc_path", acc_path)
    f.compute_kernel(X[:24], X[24:])
    K, Ki, sd = oracle_mod.run("c", X[:24], X[24:], g, m, queue, T=T, approx=True, delta=delta, max_iters=max_iters,
                               skip_variance=skip)
    Kn = oracle_mod.normalise(K, 36)
    if skip:
        assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
        assert np.array_equal(f.get_kernel_packed(), Kn)
        assert f.get_stdevs() == []
    else:
        np.testing.assert_allclose(f.get_unnormalised(np.float64), K, rtol=RTOL, atol=0)
        np.testing.assert_allclose(f.get_kernel_packed(), Kn, rtol=RTOL, atol=0)
        assert len(f.get_stdevs()) == len(sd)
        np.testing.assert_allclose(f.get_stdevs(), sd, rtol=RTOL, atol=0)
        assert f.get_stdevs()[0] == 3162.2775020544923


def test_dense_tensor_core_path_multi_tile_and_chunked(FastSK, oracle_mod):
    """acc_path 3 (K += C C^T with tcgen05 MMAs): several 128-row tiles incl. off-diagonal ones and a ragged last tile;
    long low-complexity sequences whose counts force the contraction to be cut into chunks of slots so that every fp32
    accumulator stays an exact integer; and the automatic choice on an EP300-shaped input."""
    rng = np.random.default_rng(11)
    g, m = 10, 6                                              # 4 kept characters x 2 bits: 256 k-mers
    X = random_seqs(rng, 333, 4, 60, 100)
    queue = rng.permutation(comb(g, m))[:100].astype(np.int32)   # 100 combinations: a full batch of 96 and a rest
    _, Ki, _ = oracle_mod.run("c", X[:250], X[250:], g, m, queue)
    for shape, u8 in ((0, 1), (1, 1), (2, 1), (1, 0), (2, 0)):   # auto; one tile per CTA; two tiles per CTA sharing the B operand;
        f = FastSK(g, m, combo_sequence=queue)                # byte (kind::i8, at most 255 windows per sequence) and fp16 operands
        f.set_option("gemm_shape", shape)                     # (three tile rows: the second pair's lower tile is past the last row)
        f.set_option("dense_u8", u8)
        f.compute_kernel(X[:250], X[250:])
        st = f.stats()
        assert st["acc_path"] == 3, "the cost model should pick the dense path for 256 k-mers x 333 sequences"
        assert st["dense_mode"] == (2 if u8 else 1)
        assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki), f"gemm_shape {shape}, dense_u8 {u8}"

    # the edge of the byte operands: homopolymers with exactly 255 windows (every count 255: the largest byte), then 256 (fp16)
    for length in (255 + g - 1, 256 + g - 1):
        Xh = [[1 + (i % 2)] * length for i in range(6)] + random_seqs(rng, 150, 4, 30, 60)
        _, Kh, _ = oracle_mod.run("c", Xh, [], g, m, queue[:20])
        f = FastSK(g, m, combo_sequence=queue[:20])
        f.set_option("acc_path", 3)
        f.compute_train(Xh)
        assert f.stats()["dense_mode"] == (2 if length == 255 + g - 1 else 1)
        assert np.array_equal(f.get_unnormalised().astype(np.uint64), Kh), f"length {length}"

    g, m = 8, 4
    X = random_seqs(rng, 14, 4, 900, 1100, True)             # homopolymers: counts up to 1093, products above 2^20
    queue = rng.permutation(comb(g, m))[:40].astype(np.int32)
    out = []
    for path in (2, 3):
        f = FastSK(g, m, combo_sequence=queue)
        f.set_option("acc_path", path)
        f.compute_train(X)
        out.append(f.get_unnormalised())
    assert np.array_equal(out[0], out[1])
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
    assert np.array_equal(out[1].astype(np.uint64), Ki)

    f = FastSK(12, 4, combo_sequence=queue[:4])              # 8 x 2 = 16 key bits: not eligible
    f.set_option("acc_path", 3)
    with pytest.raises(ValueError):
        f.compute_train(random_seqs(rng, 5, 4, 20, 30))


@pytest.mark.parametrize("cols", [32, 96])
@pytest.mark.parametrize("case", [CASES[0], CASES[4], CASES[5], CASES[9]], ids=lambda c: c[0])
def test_column_windows_of_the_row_path(FastSK, oracle_mod, case, cols):
    """A row of K that does not fit in shared memory is split into column windows (N > ~56 000 in production); option
    acc_cols forces small windows so that the multi-window path runs at test sizes: exact and variance mode."""
    name, ntr, nte, alpha, (lo, hi), g, m, batch, lowc = case
    rng = np.random.default_rng(5)
    n = min(ntr + nte, 16 * cols)                             # at most 16 windows
    X = random_seqs(rng, n, alpha, max(lo, g), hi, lowc)
    queue = rng.permutation(comb(g, m))[:24].astype(np.int32)
    f = FastSK(g, m, combo_sequence=queue)
    f.set_option("acc_path", 2)
    f.set_option("acc_cols", cols)
    f.compute_train(X)
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
    assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
    f = FastSK(g, m, 3, True, 0.025, 6, False, combo_sequence=queue)
    f.set_option("acc_path", 2)
    f.set_option("acc_cols", cols)
    f.compute_train(X)
    K, _, sd = oracle_mod.run("c", X, [], g, m, queue, T=3, approx=True, delta=0.025, max_iters=6)
    np.testing.assert_allclose(f.get_unnormalised(np.float64), K, rtol=RTOL, atol=0)
    np.testing.assert_allclose(f.get_stdevs(), sd, rtol=RTOL, atol=0)


def skewed_dna(rng, n, L, motif_len=12, frac=0.5):
    """SURVEY 8(d) skewed variant: first-order Markov, GC-rich sequences with one planted motif in `frac` of them, so that
    a few k-mers own very long runs (the uniform synthetic set has ~141 records in every run)."""
    P = np.array([[0.10, 0.40, 0.40, 0.10], [0.05, 0.45, 0.45, 0.05], [0.05, 0.45, 0.45, 0.05], [0.10, 0.40, 0.40, 0.10]])
    cdf = np.cumsum(P, axis=1)
    X = np.empty((n, L), dtype=np.int32)
    X[:, 0] = rng.integers(0, 4, size=n)
    u = rng.random((n, L))
    for t in range(1, L):
        X[:, t] = (u[:, t, None] > cdf[X[:, t - 1]]).sum(axis=1)
    motif = rng.integers(0, 4, size=motif_len)
    for i in np.flatnonzero(rng.random(n) < frac):
        p = int(rng.integers(0, L - motif_len + 1))
        X[i, p:p + motif_len] = motif
    return np.minimum(X, 3) + 1


@pytest.mark.parametrize("seg_fused", [1, 2], ids=["two_pass", "fused_bucket"])
def test_skewed_markov_dna_with_planted_motif(FastSK, oracle_mod, seg_fused):
    """Heavy-tailed run lengths at the C4 key shape (g=16, m=8): runs of thousands of records next to empty k-mers
    exercise the long-run paths of segment_kernel / bucket_segment_kernel and the load balancing of the accumulate."""
    rng = np.random.default_rng(1)
    X = skewed_dna(rng, 1500, 120)
    queue = np.array([0, 17, 4242, 9000, 12869], dtype=np.int32)     # incl. the first and last combination
    f = FastSK(16, 8, combo_sequence=queue, profile=True)
    f.set_option("seg_fused", seg_fused)
    f.compute_kernel(X[:1000], X[1000:])
    _, Ki, _ = oracle_mod.run("c", X[:1000].tolist(), X[1000:].tolist(), 16, 8, queue)
    assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
    st = f.stats()
    assert st["pair_updates"] > 20 * st["entries"]                # long runs: uniform data of this size has ~2 updates per entry


@pytest.mark.parametrize("shape", ["dna16_skewed", "dna8_lowcomplex", "dna5_r64_lowcomplex", "protein_r32_lowcomplex"])
def test_heavy_runs_on_the_tensor_cores(FastSK, oracle_mod, shape):
    """heavy_tau: runs longer than the threshold are filed as empty tasks and their update is done as H H^T by the tcgen05
    contraction (heavy_fill_kernel + syrk_tc_kernel).  A small forced threshold makes many runs heavy at test sizes; the
    result must equal the all-sparse build and the oracle bit for bit."""
    rng = np.random.default_rng(23)
    if shape == "dna16_skewed":
        g, m, X = 16, 8, skewed_dna(rng, 700, 100).tolist()
    elif shape == "dna8_lowcomplex":
        g, m, X = 7, 3, random_seqs(rng, 300, 4, 20, 120, True)
    elif shape == "dna5_r64_lowcomplex":
        g, m, X = 16, 4, random_seqs(rng, 150, 5, 40, 80, True)          # 36 key bits: 64-bit records
    else:
        g, m, X = 7, 3, random_seqs(rng, 200, 21, 16, 200, True)
    nc = comb(g, m)
    queue = rng.permutation(nc)[:min(nc, 12)].astype(np.int32)
    out = {}
    for tau, cap in ((-1, 0), (8, 0), (8, 64)):                     # off; on; on with a list that overflows
        f = FastSK(g, m, combo_sequence=queue, profile=True)
        f.set_option("acc_path", 2)
        f.set_option("heavy_tau", tau)
        f.set_option("heavy_cap", cap)
        f.set_option("batch", 4)
        f.compute_train(X)
        out[tau, cap] = (f.get_unnormalised(), f.stats())
    off, on, small = out[-1, 0], out[8, 0], out[8, 64]
    assert off[1]["heavy_tau"] == 0 and off[1]["heavy_runs"] == 0
    assert on[1]["heavy_tau"] == 8 and on[1]["heavy_runs"] > 0, "the test data must contain runs longer than 8 records"
    assert 0 < small[1]["heavy_runs"] <= min(on[1]["heavy_runs"], 64 * 3)   # three batches of four combinations, 64 columns each
    if shape == "dna16_skewed":
        assert small[1]["heavy_runs"] < on[1]["heavy_runs"], "the 64-column list must overflow on this input"
    assert np.array_equal(off[0], on[0]) and np.array_equal(off[0], small[0])
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
    assert np.array_equal(on[0].astype(np.uint64), Ki)


@pytest.mark.parametrize("heavy_tau,cols", [(-1, 0), (8, 0), (-1, 64), (8, 64)], ids=["plain", "heavy", "windows", "heavy_windows"])
def test_32bit_id_stream(FastSK, oracle_mod, heavy_tau, cols):
    """More than 65 000 sequences need 32-bit ids in the id stream (4 ids per 16-byte unit, the min-with-dump-word path
    of apply_unit); option ids32 forces that layout at test sizes, alone and with heavy runs and column windows."""
    rng = np.random.default_rng(77)
    g, m = 9, 4                                                  # 5 x 2 = 10 key bits
    X = random_seqs(rng, 260, 4, 20, 90, True)
    queue = rng.permutation(comb(g, m))[:16].astype(np.int32)
    f = FastSK(g, m, combo_sequence=queue, profile=True)
    f.set_option("acc_path", 2)
    f.set_option("ids32", 1)
    f.set_option("heavy_tau", heavy_tau)
    if cols:
        f.set_option("acc_cols", cols)
    f.compute_train(X)
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
    assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
    if heavy_tau > 0:
        assert f.stats()["heavy_runs"] > 0


FUSED_CASES = [
    # name, n, alphabet, len range, g, m, low complexity, batch
    ("dna_14bit", 120, 4, (40, 160), 12, 5, False, 0),          # 7 + 7 bits: 128 buckets x 128 runs
    ("dna_10bit_lowcomplex", 90, 4, (30, 120), 11, 6, True, 3),  # 5 + 5 bits, homopolymers: one group fills a run
    ("protein_15bit", 70, 21, (20, 140), 8, 5, False, 0),        # 3 x 5 bits: 8 + 7
    ("dna_16bit", 200, 4, (150, 150), 16, 8, False, 5),          # the C4 shape
    ("dna_16bit_one_seq", 1, 4, (400, 400), 12, 4, False, 0),
    ("binary_9bit", 60, 2, (40, 90), 14, 5, True, 0),            # 9 x 1 bits: 5 + 4
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_bucket_segmentation_vs_oracle(FastSK, oracle_mod, case):
    """seg_fused 2 (one onesweep pass on the high digit + bucket_segment_kernel) against the two-pass sort +
    segment_kernel (seg_fused 1) and against the oracle: same integers, same statistics."""
    name, n, alpha, (lo, hi), g, m, lowc, batch = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    X = random_seqs(rng, n, alpha, max(lo, g), hi, lowc)
    nc = comb(g, m)
    queue = rng.permutation(nc)[:min(nc, 20)].astype(np.int32)
    out, stats = [], []
    for fused in (1, 2):
        f = FastSK(g, m, combo_sequence=queue, profile=True)
        f.set_option("acc_path", 2)
        f.set_option("seg_fused", fused)
        if batch:
            f.set_option("batch", batch)
        f.compute_train(X)
        out.append(f.get_unnormalised())
        stats.append(f.stats())
    assert np.array_equal(out[0], out[1])
    for k in ("pair_updates", "entries", "runs"):
        assert stats[0][k] == stats[1][k], k
    assert stats[1]["kernel_launches"] < stats[0]["kernel_launches"]
    _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
    assert np.array_equal(out[1].astype(np.uint64), Ki)


def test_seeded_shuffle_is_reproducible_and_exact_is_order_invariant(FastSK):
    rng = np.random.default_rng(3)
    X = random_seqs(rng, 30, 4, 15, 60)
    a = FastSK(8, 4, seed=1)
    b = FastSK(8, 4, seed=2)
    a.compute_train(X)
    b.compute_train(X)
    assert not np.array_equal(a.get_queue(), b.get_queue())
    assert sorted(a.get_queue().tolist()) == list(range(comb(8, 4)))
    assert np.array_equal(a.get_unnormalised(), b.get_unnormalised())
    c = FastSK(8, 4, seed=1)
    assert np.array_equal(a.get_queue(), c.get_queue())


def test_shards_sum_to_the_whole(FastSK):
    """(e) multi-GPU contract on one device: partial kernels of the shards add up to the full build."""
    import torch
    rng = np.random.default_rng(9)
    X = random_seqs(rng, 40, 4, 30, 80)
    g, m = 10, 6
    full = FastSK(g, m, seed=7)
    full.compute_train(X)
    codes, offsets = np.concatenate([np.asarray(x, np.int32) for x in X]), np.cumsum([0] + [len(x) for x in X]).astype(np.int64)
    from fastsk_b200 import _lib
    total = None
    for world in (3,):
        for rank in range(world):
            s = FastSK(g, m, seed=7)
            s._call("fsk_set_shard", rank, world)
            s._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), len(X), 0)
            s._call("fsk_build_partial")
            part = s.partial_tensor().clone()
            total = part if total is None else total + part
    assert np.array_equal(total.cpu().numpy(), full.get_unnormalised())


def test_synthetic_c4_shape_reduced(FastSK, oracle_mod):
    """BASELINE configs[3] shape (200-bp DNA, g=16, m=8) at N the oracle finishes in seconds."""
    rng = np.random.default_rng(0)
    X = rng.integers(1, 5, size=(600, 200), dtype=np.int32)
    queue = np.random.default_rng(1).permutation(comb(16, 8))[:6].astype(np.int32)
    f = FastSK(16, 8, combo_sequence=queue, profile=True)
    f.compute_kernel(X[:480], X[480:])
    _, Ki, _ = oracle_mod.run("c", X[:480].tolist(), X[480:].tolist(), 16, 8, queue)
    assert np.array_equal(f.get_unnormalised().astype(np.uint64), Ki)
    st = f.stats()
    assert st["record_bytes"] == 4 and st["sort_passes"] == 2 and st["key_bits"] == 16
    assert st["pair_updates"] > 0 and st["ms_accumulate"] > 0


def test_full_size_properties(FastSK, oracle_mod):
    """At a size the oracle cannot reach quickly: symmetry-free invariants of the domain.
    K_ij depends only on sequences i and j, so a sub-block of a large build must equal
    an oracle run on that subset; partial kernels are additive over combinations."""
    rng = np.random.default_rng(5)
    N = 6000
    X = rng.integers(1, 5, size=(N, 200), dtype=np.int32)
    queue = np.random.default_rng(2).permutation(comb(16, 8))[:4].astype(np.int32)
    f = FastSK(16, 8, combo_sequence=queue)
    f.compute_train(X)
    Ku = f.get_unnormalised()
    idx = np.sort(rng.choice(N, size=40, replace=False))
    _, Ki, _ = oracle_mod.run("c", X[idx].tolist(), [], 16, 8, queue)
    sub = np.array([[Ku[max(a, b) * (max(a, b) + 1) // 2 + min(a, b)] for b in idx] for a in idx], dtype=np.uint64)
    assert np.array_equal(sub, oracle_mod.unpack(Ki, 40))
    # additivity: two halves of the queue
    a = FastSK(16, 8, combo_sequence=queue[:2]); a.compute_train(X)
    b = FastSK(16, 8, combo_sequence=queue[2:]); b.compute_train(X)
    assert np.array_equal(a.get_unnormalised() + b.get_unnormalised(), Ku)
    # normalised diagonal is exactly 1 and the train kernel is symmetric
    Kt = f.get_train_kernel()
    assert np.array_equal(np.diag(Kt), np.ones(N)) and np.array_equal(Kt, Kt.T)


def test_c4_full_size_submatrix_and_additivity(FastSK, oracle_mod):
    """BASELINE configs[3] at its full size (50 000 x 200 bp, g=16, m=8; the unmodified reference cannot index N > 46 341):
    a few combinations through the whole pipeline with the production batch/wave settings.  K_ij depends only on
    sequences i and j, so random sub-blocks must equal an oracle run on those sequences; partial kernels add up; the
    diagonal is the count of matching window pairs of a sequence with itself (>= windows per combination)."""
    import torch
    N, L, g, m = 50000, 200, 16, 8
    X = np.random.default_rng(0).integers(1, 5, size=(N, L), dtype=np.int32)
    queue = np.random.default_rng(3).permutation(comb(g, m))[:6].astype(np.int32)
    f = FastSK(g, m, combo_sequence=queue, distributed=False)
    f.compute_kernel(X[:40000], X[40000:])
    part = f.partial_tensor()                                 # int64 packed triangle on the device (10 GB): index it there
    rng = np.random.default_rng(11)
    idx = np.sort(np.concatenate([rng.choice(N, size=46, replace=False), [0, N - 1]]))
    idx = np.unique(idx)
    a, b = np.meshgrid(idx, idx, indexing="ij")
    hi, lo = np.maximum(a, b).astype(np.int64), np.minimum(a, b).astype(np.int64)
    flat = torch.from_numpy((hi * (hi + 1) // 2 + lo).ravel()).cuda()
    sub = part[flat].cpu().numpy().reshape(len(idx), len(idx)).astype(np.uint64)
    _, Ki, _ = oracle_mod.run("c", X[idx].tolist(), [], g, m, queue)
    assert np.array_equal(sub, oracle_mod.unpack(Ki, len(idx)))
    diag = part[torch.from_numpy(idx.astype(np.int64) * (idx.astype(np.int64) + 3) // 2).cuda()].cpu().numpy()
    assert (diag >= len(queue) * (L - g + 1)).all()
    # additivity over the combinations, checked on the same sub-block
    f2 = FastSK(g, m, combo_sequence=queue[:2], distributed=False)
    f2.compute_kernel(X[:40000], X[40000:])
    s2 = f2.partial_tensor()[flat].cpu().numpy()
    del f2
    f3 = FastSK(g, m, combo_sequence=queue[2:], distributed=False)
    f3.compute_kernel(X[:40000], X[40000:])
    s3 = f3.partial_tensor()[flat].cpu().numpy()
    assert np.array_equal((s2 + s3).reshape(sub.shape).astype(np.uint64), sub)
    # normalised outputs: unit diagonal, train block symmetric (sampled), test rows within [0, 1]
    Kt = f.get_train_kernel_tensor()
    d = torch.diagonal(Kt)
    assert bool((d == 1.0).all())
    r = torch.from_numpy(rng.choice(40000, size=64, replace=False)).cuda()
    blk = Kt[r][:, r]
    assert bool((blk == blk.T).all()) and float(blk.max()) <= 1.0 and float(blk.min()) >= 0.0


def test_errors_and_api_surface(FastSK, tmp_path):
    with pytest.raises(ValueError):
        FastSK(5, 5)
    f = FastSK(6, 2)
    with pytest.raises(ValueError):
        f.compute_kernel([[1, 2, 3, 4, 5]], [[1, 2, 3, 4, 5, 6, 7]])          # g > shortest train (fastsk.cpp:53-55)
    with pytest.raises(ValueError):
        f.compute_kernel([[1, 2, 3, 4, 5, 6, 7]], [[1, 2, 3]])                # g > shortest test (fastsk.cpp:56-58)
    with pytest.raises(RuntimeError):
        FastSK(6, 2).get_train_kernel()                                        # getter before compute
    f = FastSK(3, 1, combo_sequence=[0, 1, 2])
    f.compute_kernel([[1, 2, 1, 2, 1], [1, 1, 1, 2, 1]], [[1, 2, 1, 2, 1], [1, 1, 2, 2, 1]])
    assert f.get_train_kernel().tolist() == [[1.0, 0.6445033866354896], [0.6445033866354896, 1.0]]
    assert f.get_test_kernel().tolist() == [[1.0, 0.6445033866354896], [0.3892494720807615, 0.5853694070049635]]
    p = tmp_path / "k.txt"
    f.save_kernel(str(p))
    rows = p.read_text().strip("\n").split("\n")
    assert len(rows) == 4 and rows[0].split(" ")[1] == "2:%e" % 0.6445033866354896
    t = f.get_train_kernel_tensor()
    assert t.is_cuda and t.shape == (2, 2) and t.cpu().numpy().tolist() == f.get_train_kernel().tolist()
    f.fit(C=1.0, kernel_type="fastsk", Ytrain=[1, 0])
    assert 0.0 <= f.score("accuracy", Ytest=[1, 0]) <= 100.0     # a percentage, like fastsk.cpp:505,529


def test_reference_run_check_auc():
    """The reference's own end-to-end check (test/run_check.py:45-64): FastSK(g=10, m=6, t=1, approx=True) on EP300,
    LinearSVC + CalibratedClassifierCV on the kernel rows, AUC >= 0.9; and the exact kernel must do at least as well."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("run_check", os.path.join(root, "examples", "run_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    train, test = os.path.join(root, "data", "EP300.train.fasta"), os.path.join(root, "data", "EP300.test.fasta")
    acc, auc, _, iters = mod.run(train, test)
    assert iters >= 1 and auc >= 0.9, f"approx: AUC {auc}"
    acc_x, auc_x, _, iters_x = mod.run(train, test, g=10, m=6)
    assert iters_x == 0 and auc_x >= 0.9, f"exact: AUC {auc_x}"



def test_directory_form_with_64bit_records(FastSK):
    """More than 65 536 sequences at 16 key bits: 33 bits of key + id, so 64-bit records (and a 32-bit id stream) in the
    directory form; the register-blocked and the warp-per-rows kernels and the task-list form must agree bit for bit."""
    import hashlib
    rng = np.random.default_rng(99)
    X = rng.integers(1, 5, size=(66000, 18), dtype=np.int32)
    X[::7] = X[0]                                                          # long runs too
    queue = np.array([3, 700, 12869], dtype=np.int32)
    digests = {}
    for name, (dd, ll) in {"dir_lean": (2, 2), "dir": (2, 1), "task_lean": (1, 2), "task": (1, 1)}.items():
        f = FastSK(16, 8, combo_sequence=queue, profile=True)
        f.set_option("acc_path", 2)
        f.set_option("seg_dir", dd)
        f.set_option("seg_lean", ll)
        f.set_option("batch", 2)
        f.compute_kernel(X[:64], X[64:])
        st = f.stats()
        assert st["record_bytes"] == 8 and st["seg_mode"] == (1 if dd == 2 else 0) + (4 if ll == 2 else 0)
        digests[name] = (hashlib.sha256(f.get_unnormalised().tobytes()).hexdigest(), st["entries"], st["runs"], st["pair_updates"])
        del f
    assert len(set(digests.values())) == 1, digests


@pytest.mark.parametrize("variant", ["plain", "blocks3", "heavy", "heavy_overflow", "windows", "ids32", "heavy_ids32_windows", "no_prefetch", "unroll4"])
@pytest.mark.parametrize("shape", ["dna16_skewed", "dna16_uniform", "dna8_lowcomplex", "dna5_15bit_r64", "protein3_15bit"])
def test_directory_form_of_the_segmentation(FastSK, oracle_mod, shape, variant):
    """seg_dir = 2: no task is filed per record; the last record of every (run, block of rows) group writes one directory
    entry and the row CTA looks its windows' tasks up by key (pack_hist_kernel's wkey).  Exact and variance mode, with
    heavy runs on the tensor cores, column windows, 32-bit ids, 32- and 64-bit records, several block sizes: bit-equal to
    the oracle and to the task-list form."""
    rng = np.random.default_rng(zlib.crc32((shape + variant).encode()))
    if shape == "dna16_skewed":
        g, m, X = 16, 8, skewed_dna(rng, 700, 100).tolist()
    elif shape == "dna16_uniform":
        g, m, X = 16, 8, random_seqs(rng, 400, 4, 16, 90)
    elif shape == "dna8_lowcomplex":
        g, m, X = 7, 3, random_seqs(rng, 300, 4, 20, 120, True)
    elif shape == "dna5_15bit_r64":
        g, m, X = 9, 4, random_seqs(rng, 200, 5, 9, 80, True)                # 15 key bits
    else:
        g, m, X = 6, 3, random_seqs(rng, 200, 21, 16, 150, True)            # 3 x 5 = 15 key bits
    nc = comb(g, m)
    queue = rng.permutation(nc)[:min(nc, 10)].astype(np.int32)

    def build(seg_dir, lean=0, **mode):
        f = FastSK(g, m, combo_sequence=queue, profile=True, **mode)
        f.set_option("acc_path", 2)
        f.set_option("seg_dir", seg_dir)
        f.set_option("seg_lean", lean)
        f.set_option("batch", 4)
        f.set_option("heavy_tau", 8 if "heavy" in variant else -1)
        if variant == "heavy_overflow":
            f.set_option("heavy_cap", 64)
        if "windows" in variant:
            f.set_option("acc_cols", 64)
        if "ids32" in variant:
            f.set_option("ids32", 1)
        if variant == "blocks3":
            f.set_option("dir_blocks", 3)
        if variant == "no_prefetch":
            f.set_option("acc_prefetch", 0)
        if variant == "unroll4":
            f.set_option("acc_unroll", 4)
        f.compute_train(X)
        return f

    if "windows" in variant and len(X) > 16 * 64:
        pytest.skip("more than 16 column windows")
    d, t = build(2), build(1)
    assert d.stats()["seg_mode"] == 5 and t.stats()["seg_mode"] == 4        # + 4: the register-blocked kernel (auto)
    assert np.array_equal(d.get_unnormalised(), t.get_unnormalised())
    for dd, ll in ((2, 1), (1, 1)):                                          # the warp-per-rows kernel, both forms
        o = build(dd, ll)
        assert o.stats()["seg_mode"] == (1 if dd == 2 else 0)
        assert np.array_equal(o.get_unnormalised(), t.get_unnormalised())
        so = o.stats()
        assert (so["entries"], so["runs"], so["pair_updates"], so["heavy_runs"]) == tuple(t.stats()[k] for k in ("entries", "runs", "pair_updates", "heavy_runs"))
    if len(X) <= 1000:
        _, Ki, _ = oracle_mod.run("c", X, [], g, m, queue)
        assert np.array_equal(d.get_unnormalised().astype(np.uint64), Ki)
    sd, st = d.stats(), t.stats()
    assert (sd["entries"], sd["runs"], sd["pair_updates"]) == (st["entries"], st["runs"], st["pair_updates"])
    if "heavy" in variant:
        assert sd["heavy_runs"] == st["heavy_runs"] and (sd["heavy_runs"] > 0 or shape == "dna16_uniform")
    if variant in ("plain", "windows") and len(X) <= 1000:
        v = build(2, t=3, approx=True, delta=0.025, max_iters=4)
        K, _, sdev = oracle_mod.run("c", X, [], g, m, queue, T=3, approx=True, delta=0.025, max_iters=4)
        np.testing.assert_allclose(v.get_unnormalised(np.float64), K, rtol=RTOL, atol=0)
        # the variance statistic is one fp64 sum over all train pairs: the reference adds them one after the other, the GPU
        # block-wise; on the skewed set (terms spread over many orders of magnitude) the two orders differ by ~1e-12
        np.testing.assert_allclose(v.get_stdevs(), sdev, rtol=4e-12 if shape == "dna16_skewed" else RTOL, atol=0)


@pytest.mark.parametrize("acc_path", [2, 3], ids=["rows", "dense_tc"])
@pytest.mark.parametrize("T,max_iters,delta", [(1, 37, 0.025), (3, -1, 0.4), (4, 9, 2.0), (2, 50, 1.2)])
def test_speculated_variance_rounds_equal_the_iteration_by_iteration_build(FastSK, oracle_mod, acc_path, T, max_iters, delta):
    """Variance mode runs several consecutive iterations of every virtual stream per launch group (spec_depth) and rolls a
    stream back when its stop rule fires inside a round.  Every depth must give the same running means, the same stdevs
    and the same stopping iteration as depth 1 and as the oracle; the large deltas make streams converge in mid-round."""
    rng = np.random.default_rng(1000 + T)
    g, m = 9, 5
    X = random_seqs(rng, 40, 4, 20, 70)
    queue = rng.permutation(comb(g, m)).astype(np.int32)
    K, _, sd = oracle_mod.run("c", X[:28], X[28:], g, m, queue, T=T, approx=True, delta=delta, max_iters=max_iters)
    base = None
    for depth in (1, 2, 5, 0):
        f = FastSK(g, m, T, True, delta, max_iters, False, combo_sequence=queue)
        f.set_option("acc_path", acc_path)
        f.set_option("spec_depth", depth)
        f.compute_kernel(X[:28], X[28:])
        got = (f.get_unnormalised(np.float64), f.get_stdevs())
        assert len(got[1]) == len(sd), f"depth {depth}: stopped at another iteration"
        np.testing.assert_allclose(got[1], sd, rtol=RTOL, atol=0)
        np.testing.assert_allclose(got[0], K, rtol=RTOL, atol=0)
        if base is None:
            base = got
        else:                    # the Welford steps are the same operations in the same order at every depth
            assert np.array_equal(got[0], base[0]) and got[1] == base[1], f"depth {depth}"


@pytest.mark.parametrize("T,depth,long_seqs", [(1, 0, False), (3, 0, False), (2, 1, False), (1, 7, False), (2, 0, True)])
def test_running_means_kept_in_registers_equal_the_streamed_form(FastSK, oracle_mod, T, depth, long_seqs):
    """Tensor-core variance mode, wf_regs = 1 (default): a thread keeps its 64 cells of the running mean in registers over all
    the slots of a round and the tiles are stored strip-major; with at most 255 windows per sequence the operands are bytes
    and the accumulators int32 (dense_u8, kind::i8), otherwise fp16 / fp32.  Same Welford steps per cell as the streamed form
    (wf_regs = 0): the means must be bit-equal, the stdevs equal up to the order of one fp64 sum, all equal to the oracle.
    300 sequences: full, diagonal and ragged tiles, test rows below the training rows; long_seqs: one sequence of 400
    characters (391 windows, counts above 255 on a low-complexity stretch) keeps the fp16 operands."""
    rng = np.random.default_rng(77 + T)
    g, m = 9, 5
    X = random_seqs(rng, 300, 4, 20, 70)
    if long_seqs:
        X[5] = [1] * 300 + rng.integers(1, 5, 100).tolist()
        X[250] = [1] * 280 + rng.integers(1, 5, 40).tolist()
    queue = rng.permutation(comb(g, m)).astype(np.int32)
    K, _, sd = oracle_mod.run("c", X[:210], X[210:], g, m, queue, T=T, approx=True, delta=0.05, max_iters=23)
    got = []
    for regs, u8 in ((1, 1), (1, 0), (0, 0)):
        f = FastSK(g, m, T, True, 0.05, 23, False, combo_sequence=queue)
        f.set_option("acc_path", 3)
        f.set_option("wf_regs", regs)
        f.set_option("dense_u8", u8)
        f.set_option("spec_depth", depth)
        f.compute_kernel(X[:210], X[210:])
        assert f.stats()["dense_mode"] == (4 if regs else 0) + (2 if regs and u8 and not long_seqs else 1)
        got.append((f.get_unnormalised(np.float64), f.get_stdevs(), f.get_train_kernel(), f.get_test_kernel()))
        assert len(got[-1][1]) == len(sd)
        np.testing.assert_allclose(got[-1][1], sd, rtol=RTOL, atol=0)
        np.testing.assert_allclose(got[-1][0], K, rtol=RTOL, atol=0)
    for other in got[1:]:
        assert np.array_equal(got[0][0], other[0])
        assert np.array_equal(got[0][2], other[2]) and np.array_equal(got[0][3], other[3])
        np.testing.assert_allclose(got[0][1], other[1], rtol=RTOL, atol=0)
    assert got[0][1] == got[1][1]                          # byte and fp16 operands: the same sums in the same order


@pytest.mark.parametrize("acc_path", [2, 3], ids=["rows", "dense_tc"])
@pytest.mark.parametrize("name", ["1.1", "EP300"])
def test_full_bundled_sets_against_the_reference_fingerprints(FastSK, name, acc_path):
    """BASELINE configs[1] and [2] at FULL size (4000 / 3574 sequences, all 210 combinations): sha256 of the unnormalised
    int64 triangle and of the normalised fp64 one, the SURVEY 8c known-answer blocks and sampled cells, all produced once by
    the unmodified reference (tests/golden/make_fingerprints.py)."""
    import hashlib
    import json
    from fastsk_b200 import FastaUtility
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fingerprints.json")
    fp = json.load(open(path))[name]
    if acc_path == 3 and name == "1.1":
        pytest.skip("20 key bits: not eligible for the dense path")
    fu = FastaUtility()
    from conftest import DATA_DIR
    tr, _ = fu.read_data(os.path.join(DATA_DIR, name + ".train.fasta"))
    te, _ = fu.read_data(os.path.join(DATA_DIR, name + ".test.fasta"))
    assert (len(tr), len(te)) == (fp["n_train"], fp["n_test"])
    f = FastSK(fp["g"], fp["m"], seed=3)                         # exact mode: any order of the combinations
    f.set_option("acc_path", acc_path)
    f.compute_kernel(tr, te)
    Ki = f.get_unnormalised()
    Kn = f.get_kernel_packed()
    tri = lambda i, j: i * (i + 1) // 2 + j  # noqa: E731
    assert [[int(Ki[tri(max(a, b), min(a, b))]) for b in range(3)] for a in range(3)] == fp["kat_3x3_unnormalised"]
    assert [int(Ki[c]) for c in fp["cells"]] == fp["cells_unnormalised"]
    assert [float(Kn[c]).hex() for c in fp["cells"]] == fp["cells_normalised_hex"]
    assert int(Ki.sum()) == fp["sum_unnormalised"]
    assert hashlib.sha256(Ki.tobytes()).hexdigest() == fp["sha256_unnormalised_int64"]
    assert hashlib.sha256(Kn.tobytes()).hexdigest() == fp["sha256_normalised_f64"]


def test_linear_svm_on_the_device_matches_liblinear(FastSK):
    """fit_linear_gpu (Newton + CG on the device-resident kernel rows) solves the objective of LinearSVC(C=1): the decision
    values on the test rows agree with liblinear's, and the AUC passes the reference's acceptance bar (test/run_check.py:64)."""
    from sklearn.metrics import roc_auc_score
    from sklearn.svm import LinearSVC
    from fastsk_b200 import FastaUtility
    from conftest import DATA_DIR
    fu = FastaUtility()
    tr, ytr = fu.read_data(os.path.join(DATA_DIR, "EP300.train.fasta"))
    te, yte = fu.read_data(os.path.join(DATA_DIR, "EP300.test.fasta"))
    tr, ytr, te, yte = tr[::3], ytr[::3], te[::5], yte[::5]                 # (the files list one class after the other)
    f = FastSK(10, 6, seed=0)
    f.compute_kernel(tr, te)
    f.fit_linear_gpu(ytr, C=1.0)
    s_gpu = f.decision_function_gpu().cpu().numpy()
    svc = LinearSVC(C=1.0, tol=1e-8, max_iter=200000).fit(f.get_train_kernel(), ytr)
    s_cpu = svc.decision_function(f.get_test_kernel())
    assert np.corrcoef(s_gpu, s_cpu)[0, 1] > 0.9999
    assert np.abs(s_gpu - s_cpu).max() < 5e-3 * max(1.0, np.abs(s_cpu).max())
    auc = f.score_gpu(yte, "auc")
    assert abs(auc - roc_auc_score(yte, s_cpu)) < 1e-3 and auc >= 0.9
    assert 50.0 < f.score_gpu(yte, "accuracy") <= 100.0


def test_division_by_the_iteration_number_is_correctly_rounded():
    """The Welford step's d / iter uses a reciprocal and two FMAs (div_by_iter); it must give the bits of the IEEE division."""
    import ctypes
    from fastsk_b200 import _lib
    lib = _lib.load()
    bad = ctypes.c_uint64(123)
    for seed in (1, 2, 3):
        assert lib.fsk_selftest_division(0, seed, 200_000_000, ctypes.byref(bad)) == 0
        assert bad.value == 0
