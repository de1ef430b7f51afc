#!/bin/bash
# GPU session 17 (2 GPUs): the driver's multi-rank launch of the bench, and a 2-rank public-API job on EP300 (dense path + NCCL reduce)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 \
    > gpurun_out/s17_bench_n2.json 2> gpurun_out/s17_bench_n2.err
tail -c 1500 gpurun_out/s17_bench_n2.json | head -c 700; echo
tail -3 gpurun_out/s17_bench_n2.err
cat > /tmp/dist_ep300.py <<'PY'
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from fastsk_b200 import FastSK, FastaUtility
dist.init_process_group("nccl"); rank = dist.get_rank(); torch.cuda.set_device(rank)
root = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
fu = FastaUtility(); Xtr, _ = fu.read_data(f"{root}/data/EP300.train.fasta"); Xte, _ = fu.read_data(f"{root}/data/EP300.test.fasta")
out = {}
for mode, kw in (("exact", {}), ("approx_t20", dict(t=20, approx=True, max_iters=10))):
    f = FastSK(10, 6, seed=0, **kw); t0 = time.perf_counter(); f.compute_kernel(Xtr, Xte); dt = time.perf_counter() - t0
    out[mode] = (f.get_train_kernel(), f.stats()["acc_path"], dt, f.get_stdevs())
if rank == 0:
    for mode, kw in (("exact", {}), ("approx_t20", dict(t=20, approx=True, max_iters=10))):
        g = FastSK(10, 6, seed=0, distributed=False, **kw); g.compute_kernel(Xtr, Xte)
        K1 = g.get_train_kernel(); K2, path, dt, sd = out[mode]
        print(mode, "acc_path", path, "2-rank s", round(dt, 4), "max |diff| vs 1 rank", float(np.abs(K1 - K2).max()), "stdevs equal", np.allclose(sd, g.get_stdevs(), rtol=1e-12))
dist.destroy_process_group()
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 /tmp/dist_ep300.py 2>&1 | grep -v Warning | tail -5
