"""``FastSK``: the host-side mirror of the reference's pybind11 class.

Same constructor (positional order g, m, t, approx, delta, max_iters, skip_variance --
bindings.cpp:14-22), same method names (bindings.cpp:23-44).  The work is done by the CUDA
library behind include/fastsk_b200.h, called through ctypes; this file only flattens the
inputs and forwards calls.  Multi-GPU comes in two forms, both without a CPU path:
  * a plain script: ``FastSK(...)`` drives every visible GPU from one process (``devices="auto"``; the
    library runs one host thread per GPU -- fsk_set_devices -- like the reference runs one std::thread per
    stream, fastsk_kernel.cpp:54-94); torch is not needed;
  * under torchrun (one process per GPU): every rank builds the partial kernel of its shard, the ranks
    exchange CUDA IPC handles of their partial buffers through torch.distributed, and every rank
    normalises its share of the output rows, summing all partials over NVLink as it reads them
    (``reduce="peer"``); ``reduce="allreduce"`` is the plain NCCL form (every rank ends with everything).
If the CUDA extension is missing or no B200 is visible the calls raise.

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
import itertools
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
    if n and isinstance(X[0], list):          # lists of Python ints (the reference's input): one C-level pass, ~25 % faster
        try:
            return np.fromiter(itertools.chain.from_iterable(X), dtype=np.int32, count=int(offsets[-1])), offsets
        except TypeError:                     # a mix of lists and arrays: the general path below
            pass
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


def pinned_empty(shape, dtype=np.float64):
    """Page-locked host array owned by the library (cudaHostAlloc): getters and compute_kernel move it at PCIe speed,
    and a team of GPUs fills / reads it over all their links at once.  Freed when the array is garbage-collected."""
    lib = _lib.load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = ctypes.c_void_p()
    rc = lib.fsk_host_alloc(ctypes.byref(ptr), max(1, n * dt.itemsize))
    _lib.check(lib, None, rc)
    buf = (ctypes.c_char * max(1, n * dt.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.fsk_host_free, ptr.value)
    return arr


_BIG_OUTPUT = 64 << 20      # outputs at least this large are allocated pinned
_SHM_KEEP = []              # shared-memory segments stay mapped for the life of the process (arrays may outlive any object)


def shard_rows(n, rank, world, weights=None):
    """Rows [r0, r0 + nr) of an n-row output that rank `rank` of `world` normalises and holds (fsk_finalize, sharded): equal
    shares, or shares by the ranks' output weights (fsk_set_output_weights; the same arithmetic as the library's)."""
    import math
    w = [1.0] * world if weights is None else [float(x) for x in weights]
    total, before, mine = sum(w), sum(w[:rank]), w[rank]
    cut = lambda c: int(min(float(n), math.floor(n * (c / total) + 1e-9)))   # noqa: E731
    r0 = cut(before)
    end = n if before + mine >= total - 1e-12 else cut(before + mine)
    return r0, end - r0


_WEIGHTS = {}


def output_weights(dist):
    """COLLECTIVE.  Which ranks hand the results back to the host: all of them (None), unless the box's GPUs reach host memory
    unequally -- then only the faster half gets output rows (weights 1 / 0).  Found by timing a 256 MB device -> host copy on
    all ranks at once and on the faster half alone (fsk_probe_d2h); once per process.  FSK_EQUAL_OUTPUT_SHARES=1 skips it."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if world in _WEIGHTS:
        return _WEIGHTS[world]
    wts = None
    if world >= 4 and dist.get_backend() == "nccl" and not os.environ.get("FSK_EQUAL_OUTPUT_SHARES"):
        lib = _lib.load()
        dev, nbytes = torch.cuda.current_device(), 256 << 20

        def probe(active):
            t = ctypes.c_double(0.0)
            torch.cuda.synchronize()
            dist.barrier()
            if active:
                rc = lib.fsk_probe_d2h(dev, nbytes, ctypes.byref(t))
                if rc:
                    t.value = float("inf")
            out = torch.zeros(world, dtype=torch.float64, device="cuda")
            mine = torch.tensor([t.value], dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(out, mine)
            return out.cpu().tolist()

        probe(True)                                       # (first touch of the scratch buffers)
        t_all = probe(True)
        if all(x > 0 and x != float("inf") for x in t_all):
            order = sorted(range(world), key=lambda r: t_all[r])
            sub = set(order[:world // 2])
            t_sub = probe(rank in sub)
            agg_all = world * nbytes / max(t_all)
            agg_sub = len(sub) * nbytes / max(t_sub[r] for r in sub)
            if agg_sub > 1.25 * agg_all:
                wts = [1.0 if r in sub else 0.0 for r in range(world)]
            if rank == 0 and os.environ.get("FSK_TRACE"):
                print("[py] device -> host, all %d ranks at once: %.1f GB/s; the faster half alone: %.1f GB/s; output rows from ranks %s"
                      % (world, agg_all / 1e9, agg_sub / 1e9, sorted(sub) if wts else "all"), flush=True)
    _WEIGHTS[world] = wts
    return wts


def shared_output(rows, cols, dist=None):
    """COLLECTIVE over the ranks of a torchrun launch: a rows x cols float64 array in POSIX shared memory that every
    rank maps, with this rank's share of the rows (shard_rows) page-locked.  Pass it as ``out=`` to get_train_kernel /
    get_test_kernel: every rank then copies its rows over its own PCIe link and all ranks see the whole matrix.
    Creating it touches every page, so create it once and re-use it across computes."""
    from multiprocessing import shared_memory
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    nbytes = max(8, int(rows) * int(cols) * 8)
    box = [None]
    shm = None
    if rank == 0:
        try:
            shm = shared_memory.SharedMemory(create=True, size=nbytes)
            box = [shm.name]
        except OSError as e:
            box = ["!" + str(e)]
    dist.broadcast_object_list(box, src=0)
    if box[0].startswith("!"):
        raise MemoryError("cannot create %d bytes of shared memory for the kernel matrix: %s" % (nbytes, box[0][1:]))
    if rank != 0:
        shm = shared_memory.SharedMemory(name=box[0])
        try:                                    # the creator owns the segment: this process's tracker must not unlink it
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    arr = np.ndarray((rows, cols), dtype=np.float64, buffer=shm.buf)
    dist.barrier()
    if rank == 0:
        shm.unlink()                            # the mappings stay valid; nothing is left behind in /dev/shm
    r0, nr = shard_rows(rows, rank, world, output_weights(dist))
    if nr:
        lib = _lib.load()
        base = arr.ctypes.data
        lo = (base + r0 * cols * 8) & ~4095
        hi = min((base + (r0 + nr) * cols * 8 + 4095) & ~4095, (base + shm.size + 4095) & ~4095)
        lib.fsk_host_register(ctypes.c_void_p(lo), hi - lo)      # best effort: an unpinned range still works, only slower
    _SHM_KEEP.append(shm)
    dist.barrier()
    return arr


class FastSK:
    def __init__(self, g, m, t=-1, approx=False, delta=0.025, max_iters=-1, skip_variance=False, *,
                 seed=None, combo_sequence=None, device=None, devices="auto", distributed="auto", reduce="peer",
                 profile=False):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.fsk_create(ctypes.byref(self._h), int(g), int(m), int(t), int(bool(approx)), float(delta),
                                  int(max_iters), int(bool(skip_variance)))
        _lib.check(self._lib, None, rc)
        self.g, self.m, self.t = int(g), int(m), int(t)
        self.approx, self.delta, self.max_iters, self.skip_variance = bool(approx), float(delta), int(max_iters), bool(skip_variance)
        self._distributed = distributed
        if reduce not in ("peer", "allreduce"):
            raise ValueError("reduce must be 'peer' or 'allreduce'")
        self._reduce = reduce
        self._devices = devices if device is None else None      # an explicit device pins the handle to that GPU
        self._sharded = False       # outputs are row slices of this rank (torchrun + peer finalisation)
        self._shm = []
        self._labels = (None, None)
        self._clf = None
        self._seed, self._combo_sequence = seed, combo_sequence
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
        self._compute_flat(ctr, otr, len(otr) - 1, len(ote) - 1, cte, ote)     # (the two halves go down as they are)

    def compute_train(self, Xtrain):
        """bindings.cpp:28-31 / fastsk.cpp:120-188."""
        codes, offsets = _flatten(Xtrain)
        self._compute_flat(codes, offsets, len(offsets) - 1, 0)

    # combinations x windows above which a plain script fans out over every visible GPU (below it one GPU finishes in
    # milliseconds and seven more uploads would only add latency)
    _TEAM_WORK = 1 << 31

    def _pick_devices(self, n_windows):
        """In-process team of GPUs for a plain (non-torchrun) call."""
        dev = self._devices
        if dev is None:
            return
        if isinstance(dev, str):
            if dev not in ("auto", "all"):
                raise ValueError("devices must be 'auto', 'all' or a list of CUDA ordinals")
            if dev == "auto":
                n_comb = len(self.get_queue()) if not self.approx else max(1, self.max_iters) * (20 if self.t == -1 else self.t)
                if n_comb * max(1, n_windows) < self._TEAM_WORK:
                    return
            rc = self._lib.fsk_set_devices(self._h, None, -1)
            if rc == _lib.FSK_ECUDA:          # no device visible: the compute call reports it
                return
            _lib.check(self._lib, self._h, rc)
        else:
            arr = (ctypes.c_int * len(dev))(*[int(d) for d in dev])
            self._call("fsk_set_devices", arr, len(dev))

    def _compute_flat(self, codes, offsets, n_train, n_test, codes_test=None, offsets_test=None):
        self._clf = None
        self._sharded = False
        rank, world, dist = self._dist()
        cp, op = codes.ctypes.data_as(c_i32p), offsets.ctypes.data_as(c_i64p)
        cp2 = codes_test.ctypes.data_as(c_i32p) if codes_test is not None else None
        op2 = offsets_test.ctypes.data_as(c_i64p) if codes_test is not None else None
        if dist is None:
            self._pick_devices(len(codes) + (len(codes_test) if codes_test is not None else 0))
            self._call("fsk_compute_split", cp, op, n_train, cp2, op2, n_test)
            return
        # one process per GPU: every rank builds the partial kernel of its shard of the combinations
        # (virtual streams in variance mode)
        import time
        import torch
        trace = os.environ.get("FSK_TRACE")
        t = [time.perf_counter()]

        def lap(what):
            if trace:
                t.append(time.perf_counter())
                print("[py] rank %d %s %.3f ms" % (rank, what, (t[-1] - t[-2]) * 1e3), flush=True)

        nccl = dist.get_backend() == "nccl"
        if nccl:
            self._call("fsk_set_device", torch.cuda.current_device())
        self._agree_on_queue(dist, rank)
        self._call("fsk_release_peers")
        self._call("fsk_set_shard", rank, world)
        self._call("fsk_upload_split", cp, op, n_train, cp2, op2, n_test)
        lap("upload")
        self._call("fsk_build_partial")
        lap("build_partial")
        if nccl and self._reduce == "peer" and self._exchange_peers(dist, world):
            wts = output_weights(dist)
            if wts is not None:
                arr = (ctypes.c_double * world)(*wts)
                self._call("fsk_set_output_weights", arr, world)
            lap("exchange of the IPC handles + mapping")
            dist.barrier()                      # every rank's partial is complete before anybody reads it
            lap("barrier (slowest shard)")
            self._call("fsk_finalize")          # this rank's rows, summing all partials over NVLink
            lap("finalize (merge + normalise)")
            dist.barrier()                      # nobody resets its partial while a peer still reads it
            if not self._keep_peers:
                self._call("fsk_release_peers")
            self._sharded = True
            lap("closing barrier")
            return
        self.reduce_partial(self.partial_tensor(), dist)
        torch.cuda.current_stream().synchronize()
        self._call("fsk_finalize")

    def _agree_on_queue(self, dist, rank):
        """The ranks must walk ONE shuffled queue (their shards are slices of it): without seed= or combo_sequence=, rank 0
        draws the wall-clock seed the reference would use (fastsk_kernel.cpp:36-38) and every rank takes it."""
        if self._seed is None and self._combo_sequence is None:
            import time
            box = [int(time.time()) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self._call("fsk_set_seed", box[0])

    _keep_peers = False     # tests: keep the peers mapped so that get_unnormalised() sees the summed kernel

    def _exchange_peers(self, dist, world):
        """CUDA IPC handles of all ranks' partial buffers -> fsk_set_peer_partials.  False (on every rank) if any rank
        could not map its peers; the caller then falls back to the NCCL all-reduce."""
        import torch
        buf = ctypes.create_string_buffer(_lib.FSK_IPC_HANDLE_BYTES)
        ok = 1
        try:
            self._call("fsk_ipc_export_partial", buf)
        except RuntimeError:
            ok = 0
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        allh = torch.empty(world * _lib.FSK_IPC_HANDLE_BYTES, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        if ok:
            raw = bytes(allh.cpu().numpy().tobytes())
            try:
                self._call("fsk_set_peer_partials", raw, world)
            except RuntimeError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self._call("fsk_release_peers")
            return False
        return True

    @staticmethod
    def reduce_partial(part, dist):
        """The plain-collective form of the merge (``reduce="allreduce"``, and the gloo CPU tests): sum the ranks' partial
        kernels in place (int64 in the integer modes, float64 running means in variance mode)."""
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

    def output_rows(self):
        """(train_r0, train_nr, test_r0, test_nr): the rows of the two kernels this rank holds on its device (all rows
        unless the finalisation was sharded over the ranks of a torchrun launch)."""
        v = [ctypes.c_int64() for _ in range(4)]
        self._call("fsk_output_rows", *[ctypes.byref(x) for x in v])
        return tuple(x.value for x in v)

    def _get_kernel(self, fn, rows, out):
        n_train = self._shape()[0]
        if not self._sharded:
            if out is None:
                nbytes = rows * n_train * 8
                out = pinned_empty((rows, n_train)) if nbytes >= _BIG_OUTPUT else np.empty((rows, n_train), dtype=np.float64)
            self._call(fn, out.ctypes.data_as(c_f64p))
            return out
        # torchrun, sharded finalisation: every rank copies ITS rows, over its own PCIe link, into one buffer all ranks
        # share (POSIX shared memory created here, or the caller's `out` if all ranks were given the same memory)
        rank, world, dist = self._dist()
        if out is None:
            out = self._shared_empty((rows, n_train), dist, rank)
        self._call(fn, out.ctypes.data_as(c_f64p))
        dist.barrier()
        return out

    def _shared_empty(self, shape, dist, rank):
        return shared_output(shape[0], shape[1], dist)

    def get_train_kernel(self, out=None):
        """bindings.cpp:32 / fastsk.cpp:190-200: n_train x n_train (float64 ndarray).  Under torchrun with the sharded
        finalisation the call is COLLECTIVE (every rank calls it): each rank copies its rows into one shared-memory array all
        ranks map -- created here, or pass ``out=shared_output(...)`` (the same buffer on every rank) to re-use one."""
        return self._get_kernel("fsk_get_train_kernel", self._shape()[0], out)

    def get_test_kernel(self, out=None):
        """bindings.cpp:33 / fastsk.cpp:202-217: n_test x n_train (float64 ndarray)."""
        return self._get_kernel("fsk_get_test_kernel", self._shape()[1], out)

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
        """Device-resident train kernel as a zero-copy torch tensor (DLPack-style hand-off).  After a sharded finalisation
        (torchrun, ``reduce="peer"``) it holds this rank's rows only: ``output_rows()`` says which."""
        return self._device_kernel("fsk_train_kernel_device", 0)

    def get_test_kernel_tensor(self):
        return self._device_kernel("fsk_test_kernel_device", 1)

    def _device_kernel(self, fn, which):
        import torch
        n_train, n_test, _, _ = self._shape()
        ptr = ctypes.c_void_p()
        self._call(fn, ctypes.byref(ptr))
        tr0, trn, te0, ten = self.output_rows()
        rows = (ten if which else trn) if self._sharded else (n_test if which else n_train)
        if self.stats()["n_devices"] > 1:
            raise RuntimeError("a team of in-process GPUs holds the kernel in row shares, one per device; use device=<one GPU> "
                               "for a device-resident result")
        return torch.as_tensor(_DeviceArray(ptr.value, (rows, n_train), "<f8", self), device="cuda")

    def stats(self):
        st = _lib.FskStats()
        self._call("fsk_get_stats", ctypes.byref(st))
        return st.as_dict()

    # ------------------------------------------------------------------ learner hand-off on the device
    def fit_linear_gpu(self, Ytrain, C=1.0, tol=1e-6, max_newton=60, max_cg=200):
        """The reference's real consumer -- a linear SVM on the ROWS of the train kernel (empirical kernel map,
        test/run_check.py:48-61, examples/run.py:113-118) -- trained on the device-resident kernel, so the n_train x n_train
        matrix never crosses PCIe (16 GB at N = 50 000).  Same model as ``LinearSVC(C=C)`` (L2-regularised squared hinge,
        intercept as a regularised constant feature): minimise 1/2 |w|^2 + C sum_i max(0, 1 - y_i (K_i . w))^2 by a
        trust-region-free Newton method with conjugate gradients (every product is one fp64 GEMV over the kernel).  torch is
        used for the GEMVs only.  Needs all train rows on this handle's device (not a sharded finalisation)."""
        import torch
        if self._sharded:
            raise RuntimeError("fit_linear_gpu needs the whole train kernel on one device (reduce='allreduce' under torchrun)")
        K = self.get_train_kernel_tensor()
        n = K.shape[0]
        y = torch.as_tensor(np.where(np.asarray(Ytrain).ravel() > 0, 1.0, -1.0), dtype=torch.float64, device=K.device)
        if y.numel() != n:
            raise ValueError("Ytrain has %d labels for %d train sequences" % (y.numel(), n))
        w = torch.zeros(n + 1, dtype=torch.float64, device=K.device)          # [weights | intercept]

        def X(v):          # rows of [K | 1] times v
            return K @ v[:n] + v[n]

        def XT(u):
            return torch.cat([K.t() @ u, u.sum().reshape(1)])

        def objective(wv, z):
            return 0.5 * float(wv @ wv) + C * float(torch.clamp(1 - y * z, min=0).pow(2).sum())

        z = X(w)
        f = objective(w, z)
        for _ in range(max_newton):
            act = (1 - y * z) > 0
            a = act.to(torch.float64)
            grad = w + 2 * C * XT(a * (z - y))
            gn = float(grad.norm())
            if gn <= tol * max(1.0, float(n) ** 0.5):
                break
            # Newton direction: (I + 2C X_I^T X_I) d = -grad, by conjugate gradients
            d = torch.zeros_like(w)
            r = -grad.clone()
            p = r.clone()
            rs = float(r @ r)
            for _ in range(max_cg):
                Hp = p + 2 * C * XT(a * X(p))
                alpha = rs / float(p @ Hp)
                d += alpha * p
                r -= alpha * Hp
                rs_new = float(r @ r)
                if rs_new ** 0.5 <= 0.1 * gn:
                    break
                p = r + (rs_new / rs) * p
                rs = rs_new
            Xd = X(d)
            step, gd = 1.0, float(grad @ d)
            while True:                                  # backtracking on the exact objective
                z_new = z + step * Xd
                f_new = objective(w + step * d, z_new)
                if f_new <= f + 1e-4 * step * gd or step < 1e-10:
                    break
                step *= 0.5
            w, z = w + step * d, z_new
            if f - f_new <= 1e-12 * max(1.0, abs(f)):
                f = f_new
                break
            f = f_new
        self._w_gpu = w
        return self

    def decision_function_gpu(self, which="test"):
        """K_test . w + b (or the train rows) as a device tensor."""
        if getattr(self, "_w_gpu", None) is None:
            raise RuntimeError("decision_function_gpu called before fit_linear_gpu")
        K = self.get_test_kernel_tensor() if which == "test" else self.get_train_kernel_tensor()
        n = K.shape[1]
        return K @ self._w_gpu[:n] + self._w_gpu[n]

    def score_gpu(self, Ytest, metric="auc"):
        """Accuracy (a percentage, like fastsk.cpp:505,529) or AUC of the device-trained SVM; only n_test scores cross PCIe."""
        if metric not in ("accuracy", "auc"):
            raise ValueError("metric argument must be 'accuracy' or 'auc'")
        s = self.decision_function_gpu("test").cpu().numpy()
        yt = np.where(np.asarray(Ytest).ravel() > 0, 1, 0)
        if metric == "accuracy":
            return 100.0 * float(((s > 0).astype(int) == yt).mean())
        order = np.argsort(s, kind="mergesort")            # AUC by ranks (ties get the average rank)
        ranks = np.empty(len(s), dtype=np.float64)
        ss = s[order]
        i = 0
        while i < len(ss):
            j = i
            while j + 1 < len(ss) and ss[j + 1] == ss[i]:
                j += 1
            ranks[order[i:j + 1]] = 0.5 * (i + j) + 1
            i = j + 1
        npos, nneg = int(yt.sum()), int(len(yt) - yt.sum())
        if npos == 0 or nneg == 0:
            raise ValueError("AUC needs both classes in Ytest")
        return float((ranks[yt == 1].sum() - npos * (npos + 1) / 2) / (npos * nneg))

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
            return 100.0 * float(self._clf.score(K, self._labels[1]))      # a percentage, like fastsk.cpp:505,529
        from sklearn.metrics import roc_auc_score
        pos = list(self._clf.classes_).index(1) if 1 in self._clf.classes_ else -1
        return float(roc_auc_score(self._labels[1], self._clf.predict_proba(K)[:, pos]))
