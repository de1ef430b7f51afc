"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes
import os
import re
from math import comb

import numpy as np
import pytest

from conftest import DATA_DIR, ROOT


def test_library_loads_and_exports_every_declared_symbol():
    from fastsk_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "fastsk_b200.h")).read()
    declared = set(re.findall(r"\b(fsk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.fsk_version()


def test_stats_struct_matches_header():
    from fastsk_b200 import _lib
    header = open(os.path.join(ROOT, "include", "fastsk_b200.h")).read()
    body = header[header.index("typedef struct fsk_stats {"):header.index("} fsk_stats;")]
    names = []
    for decl in re.findall(r"(?:int64_t|int32_t|double)\s+([^;]+);", body):
        names += [n.strip() for n in decl.split(",")]
    assert names == [n for n, _ in _lib.FskStats._fields_]


def test_constructor_validation_and_queue():
    from fastsk_b200 import FastSK
    for g, m in [(3, 3), (2, 5), (0, 0)]:
        with pytest.raises(ValueError):
            FastSK(g, m)
    with pytest.raises(ValueError):
        FastSK(6, 2, t=0)
    f = FastSK(10, 6, seed=123)
    q = f.get_queue()
    assert sorted(q.tolist()) == list(range(comb(10, 6)))
    assert np.array_equal(q, FastSK(10, 6, seed=123).get_queue())
    assert not np.array_equal(q, FastSK(10, 6, seed=124).get_queue())
    assert FastSK(10, 6, combo_sequence=[5, 1, 7]).get_queue().tolist() == [5, 1, 7]
    with pytest.raises(ValueError):
        FastSK(10, 6, combo_sequence=[5, 999]).get_queue()
    with pytest.raises(ValueError):
        f.set_option("no_such_option", 1)


def test_option_validation_needs_no_device():
    """fsk_set_option checks its values on the host: the path selectors and the test hooks of the newer stages."""
    from fastsk_b200 import FastSK
    f = FastSK(10, 6)
    for key, good, bad in (("acc_path", 3, 4), ("seg_fused", 2, 3), ("heavy_tau", -1, -2), ("heavy_cap", 128, 100),
                           ("acc_cols", 64, 48), ("batch", 384, 385), ("seg_dir", 2, 3), ("seg_lean", 2, 3), ("dir_blocks", 64, 0),
                           ("spec_depth", 384, 385), ("pf_stride", 64, 96)):   # (wf_regs / dense_u8 are booleans: any value)
        f.set_option(key, good)
        with pytest.raises(ValueError):
            f.set_option(key, bad)
    st = f.stats()
    assert {"acc_path", "heavy_tau", "heavy_runs", "n_devices", "seg_mode", "dense_mode"} <= set(st) and st["heavy_runs"] == 0 and st["dense_mode"] == 0


def test_seeded_queue_equals_libstdcxx_reference_shuffle():
    """fastsk_kernel.cpp:31-38: std::shuffle over 0..C-1 with std::default_random_engine(seed)."""
    import subprocess, tempfile
    src = r"""
    #include <algorithm>
    #include <cstdio>
    #include <random>
    #include <vector>
    int main() { std::vector<int> v(35); for (int i = 0; i < 35; i++) v[i] = i;
      auto rng = std::default_random_engine{}; rng.seed(77); std::shuffle(v.begin(), v.end(), rng);
      for (int x : v) printf("%d ", x); return 0; }
    """
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.cpp"), "w").write(src)
        subprocess.run(["g++", "-O1", "-o", os.path.join(d, "s"), os.path.join(d, "s.cpp")], check=True)
        want = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()]
    from fastsk_b200 import FastSK
    assert FastSK(7, 3, seed=77).get_queue().tolist() == want


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fastsk_b200 import FastSK
    with pytest.raises(RuntimeError):
        FastSK(3, 1).compute_kernel([[1, 2, 1, 2, 1]], [[1, 2, 1, 2, 2]])


def test_argument_errors_precede_device_use():
    from fastsk_b200 import FastSK
    with pytest.raises(ValueError, match="shortest train"):
        FastSK(6, 2).compute_kernel([[1, 2, 3]], [[1, 2, 3, 4, 5, 6]])
    with pytest.raises(ValueError, match="shortest test"):
        FastSK(6, 2).compute_kernel([[1, 2, 3, 4, 5, 6]], [[1, 2]])
    with pytest.raises(ValueError):
        FastSK(6, 2).compute_kernel([[1, -2, 3, 4, 5, 6]], [[1, 2, 3, 4, 5, 6]])


def test_fasta_utility_matches_reference_encoding():
    from fastsk_b200 import FastaUtility
    r = FastaUtility()
    X, Y = r.read_data(os.path.join(DATA_DIR, "small.train.fasta"))
    assert X == [[1, 2, 1, 2, 1], [1, 1, 1, 2, 1]] and Y == [1, 0]        # ACACA / AAACA, ids from 1 in first-seen order
    X2, _ = r.read_data(os.path.join(DATA_DIR, "small.test.fasta"))
    assert X2 == [[1, 2, 1, 2, 1], [1, 1, 2, 2, 1]]
    for name, n, vocab in [("EP300", 2000, 5), ("1.1", 2339, 24), ("AImed", 1500, 56)]:
        r = FastaUtility()
        X, Y = r.read_data(os.path.join(DATA_DIR, name + ".train.fasta"))
        assert len(X) == n == len(Y) and r._vocab.size() == vocab
        codes, offsets, labels = FastaUtility().read_encoded(os.path.join(DATA_DIR, name + ".train.fasta"))
        assert labels == Y and codes.tolist() == [v for x in X for v in x] and offsets[-1] == len(codes)
    assert FastaUtility().shortest_seq(os.path.join(DATA_DIR, "small.train.fasta")) == 5


def test_read_encoded_vectorised_and_record_parsers_agree(tmp_path):
    """FASTA -> flat codes + offsets without lists of Python ints.  Plain ASCII files take one
    vectorised pass over the bytes; files with blanks, carriage returns or non-ASCII text go record by record.  Both
    must give read_data's ids (first-seen order, shared vocabulary across train and test) and labels."""
    from fastsk_b200 import FastaUtility
    for name in ("EP300", "1.1", "AImed", "small"):
        a, b = FastaUtility(), FastaUtility()
        for part in ("train", "test"):                      # one utility for both files, as every caller does
            path = os.path.join(DATA_DIR, f"{name}.{part}.fasta")
            X, Y = a.read_data(path)
            codes, offsets, labels = b.read_encoded(path)
            assert codes.dtype == np.int32 and offsets.dtype == np.int64
            assert labels == Y and codes.tolist() == [v for x in X for v in x]
            assert offsets.tolist() == np.concatenate([[0], np.cumsum([len(x) for x in X])]).tolist()
        assert str(a._vocab) == str(b._vocab)
    assert FastaUtility()._read_encoded_bytes(os.path.join(DATA_DIR, "EP300.train.fasta"), False) is not None
    assert FastaUtility()._read_encoded_bytes(os.path.join(DATA_DIR, "AImed.train.fasta"), False) is None   # text with blanks

    rng = np.random.default_rng(0)
    letters = np.array(list("ACGT"))
    lines = []
    for i in range(600):
        s = "".join(letters[rng.integers(0, 4, int(rng.integers(20, 400)))])
        if i == 500:
            s = s[:7] + "n" + s[8:]                         # a character that first appears far beyond the head
        if i % 3 == 0:
            s = s.lower()
        lines += [">%s" % ("+1" if i % 7 == 0 else ("-1" if i % 2 else "0")), s]
    for tag, text in (("lf", "\n".join(lines) + "\n"), ("no_final_newline", "\n".join(lines)), ("crlf", "\r\n".join(lines) + "\r\n")):
        path = tmp_path / f"{tag}.fasta"
        path.write_bytes(text.encode())
        a, b = FastaUtility(), FastaUtility()
        X, Y = a.read_data(str(path))
        codes, offsets, labels = b.read_encoded(str(path))
        assert labels == Y and codes.tolist() == [v for x in X for v in x] and str(a._vocab) == str(b._vocab), tag
        assert (FastaUtility()._read_encoded_bytes(str(path), False) is None) == (tag == "crlf")


def test_flatten_accepts_lists_arrays_and_flat_pairs():
    from fastsk_b200.fastsk import _flatten
    c, o = _flatten([[1, 2, 3], [4, 5]])
    assert c.tolist() == [1, 2, 3, 4, 5] and o.tolist() == [0, 3, 5]
    c, o = _flatten(np.arange(6).reshape(2, 3))
    assert c.tolist() == list(range(6)) and o.tolist() == [0, 3, 6]
    c2, o2 = _flatten((c, o))
    assert c2 is c or c2.tolist() == c.tolist()


def test_read_encoded_falls_back_for_every_byte_strip_removes(tmp_path):
    """str.strip() also removes \\x1c-\\x1f; the vectorised reader must hand such files to the record parser."""
    from fastsk_b200 import FastaUtility
    p = tmp_path / "odd.fasta"
    p.write_bytes(b">1\nAC\x1cGT\x1c\n>0\nGGTA\n")
    a, b = FastaUtility(), FastaUtility()
    X, Y = a.read_data(str(p))
    codes, offsets, labels = b.read_encoded(str(p))
    assert [codes[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)] == X
    assert list(labels) == list(Y)


def test_set_devices_validation_and_team_options():
    """fsk_set_devices needs no device to validate its list; it excludes fsk_set_shard."""
    from fastsk_b200 import FastSK, _lib
    import ctypes
    f = FastSK(8, 4)
    lib = f._lib
    arr = (ctypes.c_int * 2)(0, 0)
    assert lib.fsk_set_devices(f._h, arr, 2) == _lib.FSK_EINVAL          # listed twice
    assert lib.fsk_set_devices(f._h, None, 0) == _lib.FSK_EINVAL
    st = f.stats()
    assert st["n_devices"] == 1
    two = (ctypes.c_int * 2)(0, 1)
    for _ in range(2):                                                   # configuring a team twice (every compute does) is fine
        assert lib.fsk_set_devices(f._h, two, 2) == _lib.FSK_OK
    assert f.stats()["n_devices"] == 2
    assert lib.fsk_set_shard(f._h, 1, 2) == _lib.FSK_ESTATE              # a team shards by itself
    one = (ctypes.c_int * 1)(0)
    assert lib.fsk_set_devices(f._h, one, 1) == _lib.FSK_OK              # back to one GPU: shard 0 of 1 again
    assert f.stats()["n_devices"] == 1 and lib.fsk_set_shard(f._h, 1, 2) == _lib.FSK_OK


def test_flatten_accepts_every_input_form():
    """FastSK._flatten: lists of Python ints (one C-level pass), lists of arrays, a mix, a 2-D array, a flat pair; empty
    sequences and an empty set keep their offsets."""
    from fastsk_b200.fastsk import _flatten
    rng = np.random.default_rng(5)
    X = [rng.integers(1, 21, size=int(rng.integers(0, 30))).tolist() for _ in range(40)]
    want = np.concatenate([np.asarray(x, dtype=np.int32) for x in X])
    off = np.concatenate([[0], np.cumsum([len(x) for x in X])])
    for form in (X, [np.asarray(x, dtype=np.int64) for x in X], [x if i % 2 else np.asarray(x, dtype=np.int32) for i, x in enumerate(X)]):
        c, o = _flatten(form)
        assert c.dtype == np.int32 and o.dtype == np.int64 and np.array_equal(c, want) and np.array_equal(o, off)
    c, o = _flatten((want, off.astype(np.int64)))
    assert np.array_equal(c, want) and np.array_equal(o, off)
    A = rng.integers(1, 5, size=(7, 9))
    c, o = _flatten(A)
    assert np.array_equal(c, A.reshape(-1)) and np.array_equal(o, np.arange(8) * 9)
    c, o = _flatten([])
    assert len(c) == 0 and o.tolist() == [0]


def test_shard_rows_is_a_partition_for_any_weights():
    """shard_rows (the arithmetic fsk_finalize uses for its row shares): the ranks' ranges tile [0, n) in rank order for equal
    shares and for any non-negative weights; a rank with weight 0 holds no rows."""
    from fastsk_b200 import shard_rows
    rng = np.random.default_rng(9)
    for trial in range(200):
        world = int(rng.integers(1, 9))
        n = int(rng.integers(0, 5000))
        weights = None
        if trial % 3:
            weights = rng.integers(0, 4, size=world).astype(float).tolist()
            if sum(weights) == 0:
                weights[int(rng.integers(0, world))] = 1.0
        pos = 0
        for r in range(world):
            r0, nr = shard_rows(n, r, world, weights)
            assert r0 == pos and nr >= 0
            if weights is not None and weights[r] == 0:
                assert nr == 0
            pos += nr
        assert pos == n
    assert [shard_rows(10, r, 4, [0, 1, 0, 1]) for r in range(4)] == [(0, 0), (0, 5), (5, 0), (5, 5)]
