"""CPU oracle for the gapped k-mer kernel build -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under fastsk_b200/ imports this package.

Three independent statements of the same function:
  * ``ref``  -- oracle/_ref/libfskref.so: the unmodified reference engine driven through its
                public KernelFunction::kernel_build_parallel (fastsk_kernel.cpp:145-322) with a
                caller-supplied work queue (oracle/ref_driver.cpp).  Exists only where the
                reference sources were available at build time.
  * ``c``    -- oracle/libfsko.so: the plain-C restatement (oracle/fsk_oracle.c).
  * ``hamming_kernel`` -- NumPy, the identity K[i][j] = sum_{a in gmers(i), b in gmers(j)}
                C(g - hamming(a,b), g - m), for tiny inputs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from math import comb

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_PATH = os.path.join(_HERE, "libfsko.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libfskref.so")

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_f64p = ctypes.POINTER(ctypes.c_double)
_intp = ctypes.POINTER(ctypes.c_int)


def build(ref: bool = True) -> None:
    """Compile the oracle libraries (gcc/g++ via oracle/Makefile)."""
    subprocess.run(["make", "-C", _HERE, "liboracle"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_c = None
_ref = None


def c_lib():
    global _c
    if _c is None:
        if not os.path.exists(_C_PATH):
            build(ref=False)
        lib = ctypes.CDLL(_C_PATH)
        lib.fsko_nchoosek.restype = ctypes.c_int64
        lib.fsko_nchoosek.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.fsko_combination.restype = ctypes.c_int
        lib.fsko_combination.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _intp]
        lib.fsko_partial.restype = ctypes.c_int
        lib.fsko_partial.argtypes = [_i32p, _i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _intp, _u64p]
        lib.fsko_build.restype = ctypes.c_int
        lib.fsko_build.argtypes = [_i32p, _i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                   _i32p, ctypes.c_int, ctypes.c_int, _f64p, _u64p, _f64p, ctypes.c_int64, _i64p]
        lib.fsko_normalise.restype = None
        lib.fsko_normalise.argtypes = [_f64p, ctypes.c_int64]
        _c = lib
    return _c


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        if not ref_available():
            raise FileNotFoundError("oracle/_ref/libfskref.so not built (reference sources absent?)")
        lib = ctypes.CDLL(_REF_PATH)
        lib.fskref_build.restype = ctypes.c_int
        lib.fskref_build.argtypes = [_i32p, _i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                     _i32p, ctypes.c_int, ctypes.c_int, _f64p, _f64p, ctypes.c_int64, _i64p]
        lib.fskref_compute_kernel.restype = ctypes.c_int
        lib.fskref_compute_kernel.argtypes = [_i32p, _i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int, _f64p]
        _ref = lib
    return _ref


def flatten(X):
    """list of int sequences -> (codes int32[sum len], offsets int64[n+1])."""
    lens = np.fromiter((len(x) for x in X), dtype=np.int64, count=len(X))
    offsets = np.zeros(len(X) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    codes = np.empty(int(offsets[-1]), dtype=np.int32)
    for i, x in enumerate(X):
        codes[offsets[i]:offsets[i + 1]] = np.asarray(x, dtype=np.int32)
    return codes, offsets


def _p(a, t):
    return a.ctypes.data_as(t)


def combination(g: int, k: int, idx: int):
    pos = (ctypes.c_int * k)()
    if c_lib().fsko_combination(g, k, idx, pos):
        raise ValueError("bad combination index")
    return list(pos)


def partial(Xall, g: int, kept_positions) -> np.ndarray:
    """Integer partial kernel (packed lower triangle, uint64) of one combination."""
    codes, offsets = flatten(Xall)
    n = len(Xall)
    K = np.zeros(n * (n + 1) // 2, dtype=np.uint64)
    pos = (ctypes.c_int * len(kept_positions))(*kept_positions)
    rc = c_lib().fsko_partial(_p(codes, _i32p), _p(offsets, _i64p), n, g, len(kept_positions), pos, _p(K, _u64p))
    if rc:
        raise RuntimeError(f"fsko_partial rc={rc}")
    return K


def run(impl: str, Xtrain, Xtest, g: int, m: int, queue, T: int = 1, approx: bool = False,
        delta: float = 0.025, max_iters: int = -1, skip_variance: bool = False, normalise: bool = False):
    """Run the engine with a fixed combination queue.

    Returns (K fp64 packed lower triangle, K_int uint64 packed or None, stdevs list)."""
    Xall = list(Xtrain) + list(Xtest)
    codes, offsets = flatten(Xall)
    n = len(Xall)
    n_pairs = n * (n + 1) // 2
    q = np.ascontiguousarray(queue, dtype=np.int32)
    K = np.zeros(n_pairs, dtype=np.float64)
    cap = len(q) + 1
    sd = np.zeros(cap, dtype=np.float64)
    nsd = ctypes.c_int64(0)
    if impl == "c":
        Ki = np.zeros(n_pairs, dtype=np.uint64)
        rc = c_lib().fsko_build(_p(codes, _i32p), _p(offsets, _i64p), len(Xtrain), len(Xtest), g, m, T,
                                int(approx), delta, max_iters, int(skip_variance), _p(q, _i32p), len(q),
                                int(normalise), _p(K, _f64p), _p(Ki, _u64p), _p(sd, _f64p), cap, ctypes.byref(nsd))
        if approx and not skip_variance:
            Ki = None
    elif impl == "ref":
        Ki = None
        rc = ref_lib().fskref_build(_p(codes, _i32p), _p(offsets, _i64p), len(Xtrain), len(Xtest), g, m, T,
                                    int(approx), delta, max_iters, int(skip_variance), _p(q, _i32p), len(q),
                                    int(normalise), _p(K, _f64p), _p(sd, _f64p), cap, ctypes.byref(nsd))
    else:
        raise ValueError(impl)
    if rc:
        raise RuntimeError(f"oracle {impl} rc={rc}")
    return K, Ki, sd[:nsd.value].tolist()


def ref_compute_kernel(Xtrain, Xtest, g, m, T, approx=False, delta=0.025, max_iters=-1, skip_variance=False,
                       want_K=True):
    """The reference's own KernelFunction::compute_kernel (wall-clock-seeded shuffle)."""
    Xall = list(Xtrain) + list(Xtest)
    codes, offsets = flatten(Xall)
    n = len(Xall)
    K = np.zeros(n * (n + 1) // 2, dtype=np.float64) if want_K else None
    rc = ref_lib().fskref_compute_kernel(_p(codes, _i32p), _p(offsets, _i64p), len(Xtrain), len(Xtest), g, m, T,
                                         int(approx), delta, max_iters, int(skip_variance),
                                         _p(K, _f64p) if want_K else None)
    if rc:
        raise RuntimeError(f"fskref_compute_kernel rc={rc}")
    return K


def normalise(K_packed: np.ndarray, n: int) -> np.ndarray:
    K = np.array(K_packed, dtype=np.float64, copy=True)
    c_lib().fsko_normalise(_p(K, _f64p), n)
    return K


def unpack(K_packed: np.ndarray, n: int) -> np.ndarray:
    """packed lower triangle -> symmetric n x n."""
    out = np.zeros((n, n), dtype=K_packed.dtype)
    il = np.tril_indices(n)
    out[il] = K_packed
    out.T[il] = K_packed
    return out


def hamming_kernel(Xall, g: int, m: int) -> np.ndarray:
    """Independent restatement for tiny inputs: exact-mode unnormalised K as int64 n x n."""
    k = g - m
    w = np.array([comb(g - d, k) if g - d >= k else 0 for d in range(g + 1)], dtype=np.int64)
    gm = []
    for x in Xall:
        x = np.asarray(x, dtype=np.int64)
        gm.append(np.lib.stride_tricks.sliding_window_view(x, g) if len(x) >= g else np.zeros((0, g), np.int64))
    n = len(Xall)
    K = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for j in range(i + 1):
            d = (gm[i][:, None, :] != gm[j][None, :, :]).sum(-1)
            K[i, j] = K[j, i] = w[d].sum()
    return K
