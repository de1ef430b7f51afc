"""``FastSK``: the host-side mirror of the reference's pybind11 class.

Same constructor (positional order g, m, t, approx, delta, max_iters, skip_variance --
bindings.cpp:14-22), same method names (bindings.cpp:23-44).  The work is done by the CUDA
library behind include/fastsk_b200.h, called through ctypes; this file only flattens the
inputs, forwards calls and, when launched under torchrun, sums the per-GPU partial kernels with
one NCCL all-reduce.  If the CUDA extension is missing or no B200 is visible the calls raise;
there is no CPU path.

Deliberate deviations from the reference:
  * getters return NumPy arrays (``.tolist()`` gives the reference's list of lists) -- a Python
    list of 2.5e9 floats is not an option at N = 50 000;
  * argument errors raise ``ValueError`` where the reference prints and calls ``exit(1)``;
  * the combination order is reproducible: ``seed=`` or ``combo_sequence=`` (the reference seeds
    its shuffle with the wall clock, fastsk_kernel.cpp:36-38);
  * ``fit``/``score`` take the labels explicitly (the reference reads label arrays that nothing
    ever sets and crashes, fastsk.hpp:43-44) and run scikit-learn's libsvm on the kernel.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from ._lib import c_f64p, c_i32p, c_i64p


def _flatten(X):
    """Sequences -> (codes int32, offsets int64).  Accepts list of int lists, a 2-D integer array
    (equal lengths) or an already flat (codes, offsets) pair."""
    if isinstance(X, tuple) and len(X) == 2 and isinstance(X[0], np.ndarray):
        codes, offsets = X
        return np.ascontiguousarray(codes, dtype=np.int32), np.ascontiguousarray(offsets, dtype=np.int64)
    if isinstance(X, np.ndarray) and X.ndim == 2:
        n, L = X.shape
        return np.ascontiguousarray(X, dtype=np.int32).reshape(-1), np.arange(n + 1, dtype=np.int64) * L
    n = len(X)
    lens = np.fromiter((len(x) for x in X), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    codes = np.empty(int(offsets[-1]), dtype=np.int32)
    for i, x in enumerate(X):
        codes[offsets[i]:offsets[i + 1]] = x
    return codes, offsets


class _DeviceArray:
    """Zero-copy view of a device buffer owned by the library (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class FastSK:
    def __init__(self, g, m, t=-1, approx=False, delta=0.025, max_iters=-1, skip_variance=False, *,
                 seed=None, combo_sequence=None, device=None, distributed="auto", profile=False):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.fsk_create(ctypes.byref(self._h), int(g), int(m), int(t), int(bool(approx)), float(delta),
                                  int(max_iters), int(bool(skip_variance)))
        _lib.check(self._lib, None, rc)
        self.g, self.m, self.t = int(g), int(m), int(t)
        self.approx, self.delta, self.max_iters, self.skip_variance = bool(approx), float(delta), int(max_iters), bool(skip_variance)
        self._distributed = distributed
        self._labels = (None, None)
        self._clf = None
        if seed is not None:
            self._call("fsk_set_seed", int(seed))
        if combo_sequence is not None:
            q = np.ascontiguousarray(combo_sequence, dtype=np.int32)
            self._call("fsk_set_combo_sequence", q.ctypes.data_as(c_i32p), len(q))
        if device is not None:
            self._call("fsk_set_device", int(device))
        if profile:
            self.set_option("profile", 1)

    # ------------------------------------------------------------------ plumbing
    def _call(self, name, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        _lib.check(self._lib, self._h, rc)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.fsk_destroy(h)

    def set_option(self, key, value):
        self._call("fsk_set_option", key.encode(), int(value))

    def _dist(self):
        """(rank, world, torch.distributed or None) of the launch this process belongs to."""
        if self._distributed is False:
            return 0, 1, None
        try:
            import torch.distributed as dist
        except ImportError:
            return 0, 1, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), dist
        return 0, 1, None

    # ------------------------------------------------------------------ reference API
    def compute_kernel(self, Xtrain, Xtest):
        """bindings.cpp:23-27 / fastsk.cpp:30-118.  Blocks until the normalised kernels are on the device."""
        ctr, otr = _flatten(Xtrain)
        cte, ote = _flatten(Xtest)
        codes = np.concatenate([ctr, cte])
        offsets = np.concatenate([otr, ote[1:] + otr[-1]])
        self._compute_flat(codes, offsets, len(otr) - 1, len(ote) - 1)

    def compute_train(self, Xtrain):
        """bindings.cpp:28-31 / fastsk.cpp:120-188."""
        codes, offsets = _flatten(Xtrain)
        self._compute_flat(codes, offsets, len(offsets) - 1, 0)

    def _compute_flat(self, codes, offsets, n_train, n_test):
        self._clf = None
        rank, world, dist = self._dist()
        cp, op = codes.ctypes.data_as(c_i32p), offsets.ctypes.data_as(c_i64p)
        if dist is None:
            self._call("fsk_compute", cp, op, n_train, n_test)
            return
        # one process per GPU: every rank builds the partial kernel of its shard of the combinations
        # (virtual streams in variance mode); one NCCL all-reduce over NVLink combines them.
        import torch
        if dist.get_backend() == "nccl":
            self._call("fsk_set_device", torch.cuda.current_device())
        self._call("fsk_set_shard", rank, world)
        self._call("fsk_upload", cp, op, n_train, n_test)
        self._call("fsk_build_partial")
        self.reduce_partial(self.partial_tensor(), dist)
        torch.cuda.current_stream().synchronize()
        self._call("fsk_finalize")

    @staticmethod
    def reduce_partial(part, dist):
        """The one collective of the path: sum the ranks' partial kernels in place (int64 in the integer
        modes, float64 running means in variance mode).  NCCL over NVLink on the GPU box, gloo in CPU tests."""
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        return part

    def set_shard(self, rank, world):
        self._call("fsk_set_shard", int(rank), int(world))

    def get_shard_work(self):
        """Combination numbers (integer modes) or virtual stream ids (variance mode) of this handle's shard."""
        n = ctypes.c_int64()
        self._call("fsk_get_shard_work", None, 0, ctypes.byref(n))
        out = np.zeros(max(n.value, 1), dtype=np.int32)
        self._call("fsk_get_shard_work", out.ctypes.data_as(c_i32p), n.value, ctypes.byref(n))
        return out[:n.value]

    def partial_tensor(self):
        """This rank's unnormalised partial kernel (packed lower triangle) as a zero-copy torch tensor."""
        import torch
        ptr, n, dt = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        self._call("fsk_partial_buffer", ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(dt))
        view = _DeviceArray(ptr.value, (n.value,), "<i8" if dt.value == _lib.FSK_DT_I64 else "<f8", self)
        return torch.as_tensor(view, device=f"cuda:{torch.cuda.current_device()}")

    def _shape(self):
        a, b, c, d = (ctypes.c_int64() for _ in range(4))
        self._call("fsk_shape", ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
        return a.value, b.value, c.value, d.value

    def get_train_kernel(self, out=None):
        """bindings.cpp:32 / fastsk.cpp:190-200: n_train x n_train (float64 ndarray)."""
        n_train, _, _, _ = self._shape()
        if out is None:
            out = np.empty((n_train, n_train), dtype=np.float64)
        self._call("fsk_get_train_kernel", out.ctypes.data_as(c_f64p))
        return out

    def get_test_kernel(self, out=None):
        """bindings.cpp:33 / fastsk.cpp:202-217: n_test x n_train (float64 ndarray)."""
        n_train, n_test, _, _ = self._shape()
        if out is None:
            out = np.empty((n_test, n_train), dtype=np.float64)
        self._call("fsk_get_test_kernel", out.ctypes.data_as(c_f64p))
        return out

    def get_stdevs(self):
        """bindings.cpp:34 / fastsk.cpp:219-221."""
        n = ctypes.c_int64()
        self._call("fsk_get_stdevs", None, 0, ctypes.byref(n))
        out = np.zeros(max(n.value, 1), dtype=np.float64)
        self._call("fsk_get_stdevs", out.ctypes.data_as(c_f64p), n.value, ctypes.byref(n))
        return out[:n.value].tolist()

    def save_kernel(self, kernel_file):
        """bindings.cpp:35 / fastsk.cpp:223-237."""
        self._call("fsk_save_kernel", os.fsencode(kernel_file))

    # ------------------------------------------------------------------ beyond the reference
    def get_kernel_packed(self):
        """The reference's internal ``double* K``: normalised packed lower triangle."""
        n_train, n_test, _, _ = self._shape()
        n = n_train + n_test
        out = np.empty(n * (n + 1) // 2, dtype=np.float64)
        self._call("fsk_get_kernel_packed", out.ctypes.data_as(c_f64p))
        return out

    def get_unnormalised(self, dtype=np.int64):
        """Packed lower triangle before normalisation (int64 in the integer modes, float64 always)."""
        n_train, n_test, _, _ = self._shape()
        n = n_train + n_test
        if np.dtype(dtype) == np.int64:
            out = np.empty(n * (n + 1) // 2, dtype=np.int64)
            self._call("fsk_get_unnormalised_i64", out.ctypes.data_as(c_i64p))
        else:
            out = np.empty(n * (n + 1) // 2, dtype=np.float64)
            self._call("fsk_get_unnormalised_f64", out.ctypes.data_as(c_f64p))
        return out

    def get_queue(self):
        n = ctypes.c_int64()
        self._call("fsk_get_queue", None, 0, ctypes.byref(n))
        out = np.zeros(max(n.value, 1), dtype=np.int32)
        self._call("fsk_get_queue", out.ctypes.data_as(c_i32p), n.value, ctypes.byref(n))
        return out[:n.value]

    def get_train_kernel_tensor(self):
        """Device-resident train kernel as a zero-copy torch tensor (DLPack-style hand-off)."""
        return self._device_kernel("fsk_train_kernel_device", 0)

    def get_test_kernel_tensor(self):
        return self._device_kernel("fsk_test_kernel_device", 1)

    def _device_kernel(self, fn, which):
        import torch
        n_train, n_test, _, _ = self._shape()
        ptr = ctypes.c_void_p()
        self._call(fn, ctypes.byref(ptr))
        rows = n_test if which else n_train
        return torch.as_tensor(_DeviceArray(ptr.value, (rows, n_train), "<f8", self), device="cuda")

    def stats(self):
        st = _lib.FskStats()
        self._call("fsk_get_stats", ctypes.byref(st))
        return st.as_dict()

    # ------------------------------------------------------------------ learner hand-off
    def set_labels(self, Ytrain, Ytest=None):
        self._labels = (None if Ytrain is None else np.asarray(Ytrain).ravel(),
                        None if Ytest is None else np.asarray(Ytest).ravel())

    def fit(self, C=1.0, nu=0.5, eps=0.001, kernel_type="linear", Ytrain=None):
        """bindings.cpp:36-41 / fastsk.cpp:239-300: C-SVC with probability estimates on the train kernel;
        'linear' / 'rbf' treat kernel rows as feature vectors (gamma = 1/nfeat), 'fastsk' uses the kernel
        itself (precomputed).  LIBSVM comes from scikit-learn instead of the vendored copy."""
        from sklearn.svm import SVC
        if kernel_type not in ("linear", "fastsk", "rbf"):
            raise ValueError("kernel must be: 'linear', 'fastsk', or 'rbf'")
        if Ytrain is not None:
            self._labels = (np.asarray(Ytrain).ravel(), self._labels[1])
        if self._labels[0] is None:
            raise ValueError("fit needs the train labels: fit(..., Ytrain=...) or set_labels(Ytrain, Ytest)")
        _, _, nfeat, _ = self._shape()
        K = self.get_train_kernel()
        kern = {"linear": "linear", "fastsk": "precomputed", "rbf": "rbf"}[kernel_type]
        self._clf = SVC(C=C, kernel=kern, gamma=1.0 / max(nfeat, 1), tol=eps, probability=True, cache_size=100, shrinking=True)
        self._clf.fit(K, self._labels[0])
        return self

    def score(self, metric="auc", Ytest=None):
        """bindings.cpp:42-44 / fastsk.cpp:418-530: accuracy or AUC of the fitted SVM on the test kernel rows."""
        if metric not in ("accuracy", "auc"):
            raise ValueError("metric argument must be 'accuracy' or 'auc'")
        if self._clf is None:
            raise RuntimeError("score called before fit")
        if Ytest is not None:
            self._labels = (self._labels[0], np.asarray(Ytest).ravel())
        if self._labels[1] is None:
            raise ValueError("score needs the test labels: score(..., Ytest=...) or set_labels(Ytrain, Ytest)")
        K = self.get_test_kernel()
        if metric == "accuracy":
            return float(self._clf.score(K, self._labels[1]))
        from sklearn.metrics import roc_auc_score
        pos = list(self._clf.classes_).index(1) if 1 in self._clf.classes_ else -1
        return float(roc_auc_score(self._labels[1], self._clf.predict_proba(K)[:, pos]))
