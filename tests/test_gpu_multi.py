"""More than one rank on the CUDA path (-m gpu).

* two shards on ONE GPU, finalised through fsk_set_peer_pointers: the sharded normalisation (every rank sums all partial
  kernels as it reads them and holds only its rows) against the one-handle build -- runs on the single-GPU box too;
* the in-process team (fsk_set_devices) and the torchrun path (CUDA IPC peer finalisation, and the NCCL all-reduce form)
  against the one-GPU build: bit-equal in the integer modes, rtol 1e-12 in variance mode -- need >= 2 visible GPUs."""
import ctypes
import os
import socket
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-12

MODES = {
    "exact": dict(),
    "skip_variance": dict(t=3, approx=True, max_iters=9, skip_variance=True),
    "variance": dict(t=5, approx=True, max_iters=6),
}


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def make_inputs(seed=11, n=90, g=9, m=4, alpha=4):
    rng = np.random.default_rng(seed)
    X = [rng.integers(1, alpha + 1, size=int(rng.integers(g, 120))).tolist() for _ in range(n)]
    return X[:60], X[60:], g, m, rng.permutation(comb(g, m)).astype(np.int32)


def one_gpu(FastSK, Xtr, Xte, g, m, queue, **kw):
    f = FastSK(g, m, combo_sequence=queue, device=0, distributed=False, **kw)
    f.compute_kernel(Xtr, Xte)
    return f


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("world,weights", [(2, None), (3, None), (3, [2.0, 0.0, 1.0]), (4, [0.0, 1.0, 0.0, 1.0])])
def test_shards_on_one_gpu_sharded_finalise(mode, world, weights):
    from fastsk_b200 import FastSK, shard_rows
    from fastsk_b200.fastsk import _flatten
    Xtr, Xte, g, m, queue = make_inputs()
    ref = one_gpu(FastSK, Xtr, Xte, g, m, queue, **MODES[mode])
    ctr, otr = _flatten(Xtr)
    cte, ote = _flatten(Xte)
    codes = np.concatenate([ctr, cte])
    offsets = np.concatenate([otr, ote[1:] + otr[-1]])
    hs = []
    for r in range(world):
        f = FastSK(g, m, combo_sequence=queue, device=0, distributed=False, **MODES[mode])
        f.set_shard(r, world)
        f._call("fsk_upload", codes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                len(Xtr), len(Xte))
        f._call("fsk_build_partial")
        hs.append(f)
    ptrs = (ctypes.c_void_p * world)()
    for r, f in enumerate(hs):
        p, n, dt = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        f._call("fsk_partial_buffer", ctypes.byref(p), ctypes.byref(n), ctypes.byref(dt))
        ptrs[r] = p.value
    tr = np.full((len(Xtr), len(Xtr)), np.nan)
    te = np.full((len(Xte), len(Xtr)), np.nan)
    rows = []
    for f in hs:
        f._call("fsk_set_peer_pointers", ptrs, world)
        if weights is not None:                  # unequal shares of the output rows (a rank with weight 0 hands nothing back)
            f._call("fsk_set_output_weights", (ctypes.c_double * world)(*weights), world)
        f._call("fsk_finalize")
        rows.append(f.output_rows())
        r = len(rows) - 1
        assert (rows[r][0], rows[r][1]) == shard_rows(len(Xtr), r, world, weights)
        assert (rows[r][2], rows[r][3]) == shard_rows(len(Xte), r, world, weights)
        f.get_train_kernel(out=tr)
        f.get_test_kernel(out=te)
    assert sum(r[1] for r in rows) == len(Xtr) and sum(r[3] for r in rows) == len(Xte)     # the shares partition the rows
    if mode == "variance":
        np.testing.assert_allclose(tr, ref.get_train_kernel(), rtol=RTOL, atol=0)
        np.testing.assert_allclose(te, ref.get_test_kernel(), rtol=RTOL, atol=0)
        np.testing.assert_allclose(hs[0].get_kernel_packed(), ref.get_kernel_packed(), rtol=RTOL, atol=0)
        np.testing.assert_allclose(hs[0].get_stdevs(), ref.get_stdevs(), rtol=RTOL, atol=0)
    else:
        assert np.array_equal(tr, ref.get_train_kernel()) and np.array_equal(te, ref.get_test_kernel())
        assert np.array_equal(hs[1].get_unnormalised(), ref.get_unnormalised())      # the getters sum the reachable partials
        assert np.array_equal(hs[0].get_kernel_packed(), ref.get_kernel_packed())
    for f in hs:
        f._call("fsk_release_peers")


@pytest.mark.parametrize("mode", list(MODES))
def test_team_in_process_matches_one_gpu(mode, tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from fastsk_b200 import FastSK
    Xtr, Xte, g, m, queue = make_inputs(seed=12)
    ref = one_gpu(FastSK, Xtr, Xte, g, m, queue, **MODES[mode])
    f = FastSK(g, m, combo_sequence=queue, devices=list(range(min(n_gpus(), 4))), distributed=False, **MODES[mode])
    f.compute_kernel(Xtr, Xte)
    st = f.stats()
    assert st["n_devices"] == min(n_gpus(), 4)
    if mode == "variance":
        np.testing.assert_allclose(f.get_train_kernel(), ref.get_train_kernel(), rtol=RTOL, atol=0)
        np.testing.assert_allclose(f.get_test_kernel(), ref.get_test_kernel(), rtol=RTOL, atol=0)
        np.testing.assert_allclose(f.get_stdevs(), ref.get_stdevs(), rtol=RTOL, atol=0)
        assert len(f.get_stdevs()) == len(ref.get_stdevs())
    else:
        assert st["combos_done"] == ref.stats()["combos_done"]
        assert np.array_equal(f.get_unnormalised(), ref.get_unnormalised())
        assert np.array_equal(f.get_train_kernel(), ref.get_train_kernel())
        assert np.array_equal(f.get_test_kernel(), ref.get_test_kernel())
        assert np.array_equal(f.get_kernel_packed(), ref.get_kernel_packed())
    f.save_kernel(str(tmp_path / "k.txt"))
    ref.save_kernel(str(tmp_path / "k1.txt"))
    assert open(tmp_path / "k.txt").read() == open(tmp_path / "k1.txt").read()
    # a second compute on the same object, unseeded: the members must still agree on one queue
    h = FastSK(g, m, devices=[0, 1], distributed=False)
    h.compute_kernel(Xte, Xtr)
    h.compute_kernel(Xtr, Xte)
    e = FastSK(g, m, device=0, distributed=False)
    e.compute_kernel(Xtr, Xte)
    assert np.array_equal(h.get_unnormalised(), e.get_unnormalised())       # exact mode: any order, same integers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from fastsk_b200 import FastSK
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        Xtr, Xte, g, m, queue = make_inputs(seed=13)
        res = {}
        for mode, kw in MODES.items():
            for reduce in ("peer", "allreduce"):
                f = FastSK(g, m, combo_sequence=queue, reduce=reduce, **kw)
                f.compute_kernel(Xtr, Xte)
                tr, te = f.get_train_kernel(), f.get_test_kernel()
                if rank == 0:
                    res[f"{mode}_{reduce}_train"] = np.array(tr)
                    res[f"{mode}_{reduce}_test"] = np.array(te)
                    res[f"{mode}_{reduce}_sharded"] = np.array([f._sharded])
                    res[f"{mode}_{reduce}_stdevs"] = np.array(f.get_stdevs())
                f2 = f       # a second compute on the same object (fsk_set_device after upload used to raise)
                f2.compute_kernel(Xtr, Xte)
                if rank == 0:
                    res[f"{mode}_{reduce}_train2"] = np.array(f2.get_train_kernel())
                else:
                    f2.get_train_kernel()
                del f, f2
        # unseeded exact build: the ranks agree on one queue
        f = FastSK(g, m)
        f.compute_kernel(Xtr, Xte)
        tr = f.get_train_kernel()
        if rank == 0:
            res["unseeded_train"] = np.array(tr)
            np.savez(os.path.join(out_dir, "res.npz"), **res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_one_rank(tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from fastsk_b200 import FastSK
    mp.spawn(_rank_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = np.load(tmp_path / "res.npz")
    Xtr, Xte, g, m, queue = make_inputs(seed=13)
    for mode, kw in MODES.items():
        ref = one_gpu(FastSK, Xtr, Xte, g, m, queue, **kw)
        for reduce in ("peer", "allreduce"):
            tr, te = res[f"{mode}_{reduce}_train"], res[f"{mode}_{reduce}_test"]
            if mode == "variance":
                np.testing.assert_allclose(tr, ref.get_train_kernel(), rtol=RTOL, atol=0)
                np.testing.assert_allclose(te, ref.get_test_kernel(), rtol=RTOL, atol=0)
                np.testing.assert_allclose(res[f"{mode}_{reduce}_stdevs"], ref.get_stdevs(), rtol=RTOL, atol=0)
            else:
                assert np.array_equal(tr, ref.get_train_kernel()), (mode, reduce)
                assert np.array_equal(te, ref.get_test_kernel()), (mode, reduce)
            assert np.array_equal(res[f"{mode}_{reduce}_train2"], tr)
        assert bool(res[f"{mode}_peer_sharded"][0])          # the IPC path was really taken
    e = FastSK(g, m, device=0, distributed=False)
    e.compute_kernel(Xtr, Xte)
    assert np.array_equal(res["unseeded_train"], e.get_train_kernel())
