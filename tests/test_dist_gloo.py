"""(e) multi-GPU contract on CPU: world_size-2 gloo run of the host-side sharding + the one collective.

Each rank asks the C ABI for its shard (fsk_get_shard_work needs no device), stands in for the CUDA
partial build with the oracle, and sums the partial kernels with the same reduce_partial() the GPU path
uses.  The result must equal the single-process oracle."""
import os
import socket
from math import comb

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import oracle
    from fastsk_b200 import FastSK
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(4)
        g, m = 8, 4
        X = [rng.integers(1, 5, size=int(rng.integers(g, 50))).tolist() for _ in range(20)]
        queue = rng.permutation(comb(g, m)).astype(np.int32)
        n, n_pairs = len(X), len(X) * (len(X) + 1) // 2
        kw = {"exact": dict(t=3), "skipvar": dict(t=3, approx=True, max_iters=7, skip_variance=True),
              "variance": dict(t=5, approx=True, max_iters=6)}[mode]
        f = FastSK(g, m, combo_sequence=queue, distributed=False, **kw)
        f.set_shard(rank, world)
        work = f.get_shard_work()
        if mode == "variance":
            part = np.zeros(n_pairs, dtype=np.float64)
            for tid in work:        # stream tid alone == the engine with T = 1 on queue[tid::T]
                K, _, _ = oracle.run("c", X[:14], X[14:], g, m, queue[tid::5], T=1, approx=True, max_iters=6)
                part += K
            t = torch.from_numpy(part)
        else:
            part = np.zeros(n_pairs, dtype=np.uint64)
            for c in work:
                part += oracle.partial(X, g, oracle.combination(g, g - m, int(c)))
            t = torch.from_numpy(part.astype(np.int64))
        FastSK.reduce_partial(t, dist)
        if rank == 0:
            np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
            np.save(os.path.join(out_dir, "work0.npy"), work)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["exact", "skipvar", "variance"])
def test_world2_shards_reduce_to_oracle(tmp_path, mode, oracle_mod):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sum.npy")
    rng = np.random.default_rng(4)
    g, m = 8, 4
    X = [rng.integers(1, 5, size=int(rng.integers(g, 50))).tolist() for _ in range(20)]
    queue = rng.permutation(comb(g, m)).astype(np.int32)
    if mode == "exact":
        _, Ki, _ = oracle_mod.run("c", X[:14], X[14:], g, m, queue, T=3)
        assert np.array_equal(got.astype(np.uint64), Ki)
        assert len(np.load(tmp_path / "work0.npy")) == (comb(g, m) + 1) // 2
    elif mode == "skipvar":
        _, Ki, _ = oracle_mod.run("c", X[:14], X[14:], g, m, queue, T=3, approx=True, max_iters=7, skip_variance=True)
        assert np.array_equal(got.astype(np.uint64), Ki)
    else:
        K, _, _ = oracle_mod.run("c", X[:14], X[14:], g, m, queue, T=5, approx=True, max_iters=6)
        np.testing.assert_allclose(got, K, rtol=1e-14, atol=0)
        assert np.load(tmp_path / "work0.npy").tolist() == [0, 2, 4]


def test_shard_work_partitions_the_queue():
    from fastsk_b200 import FastSK
    q = np.random.default_rng(0).permutation(comb(9, 4)).astype(np.int32)
    for world in (1, 2, 4, 8):
        parts = []
        for r in range(world):
            f = FastSK(9, 4, combo_sequence=q, distributed=False)
            f.set_shard(r, world)
            parts.append(f.get_shard_work())
        assert sorted(np.concatenate(parts).tolist()) == sorted(q.tolist())
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # approx + skip_variance, t=4, max_iters=3 -> exactly the first 12 queue items (fastsk_kernel.cpp:257-262,275-278)
    f = FastSK(9, 4, 4, True, 0.025, 3, True, combo_sequence=q, distributed=False)
    assert sorted(f.get_shard_work().tolist()) == sorted(q[:12].tolist())
    # variance mode, t=-1 -> 20 virtual streams (fastsk_kernel.cpp:54-60), dealt round-robin
    f = FastSK(9, 4, -1, True, combo_sequence=q, distributed=False)
    f.set_shard(1, 8)
    assert f.get_shard_work().tolist() == [1, 9, 17]


def _seed_worker(rank, world, port, out_dir):
    import sys, time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from fastsk_b200 import FastSK
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        time.sleep(1.2 * rank)                    # the ranks reach the call in different wall-clock seconds
        f = FastSK(8, 4)                          # no seed=, no combo_sequence=
        f._agree_on_queue(dist, rank)
        f.set_shard(rank, world)
        np.save(os.path.join(out_dir, f"work{rank}.npy"), f.get_shard_work())
    finally:
        dist.destroy_process_group()


def test_unseeded_ranks_agree_on_one_queue(tmp_path):
    """without seed= the ranks used to shuffle with their own time(0); the shards were not a partition."""
    import torch.multiprocessing as mp
    mp.spawn(_seed_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    w0, w1 = np.load(tmp_path / "work0.npy"), np.load(tmp_path / "work1.npy")
    assert sorted(np.concatenate([w0, w1]).tolist()) == list(range(comb(8, 4)))
