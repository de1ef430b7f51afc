"""The full configs[3] build as ONE job under torchrun, with the phases of every rank on stderr (FSK_TRACE=1)."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic, queue_order, parity_sample, parity_check, N_SEQ, N_TRAIN, G, M
from fastsk_b200 import FastSK
from fastsk_b200.fastsk import pinned_empty, shared_output

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
X = synthetic()
q = queue_order()
n_test = N_SEQ - N_TRAIN
Xtr, Xte = pinned_empty((N_TRAIN, 200), np.int32), pinned_empty((n_test, 200), np.int32)
Xtr[:], Xte[:] = X[:N_TRAIN], X[N_TRAIN:]
otr, ote = shared_output(N_TRAIN, N_TRAIN, dist), shared_output(n_test, N_TRAIN, dist)
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    f = FastSK(G, M, combo_sequence=q if rep else q[:world * 8])
    f.compute_kernel(Xtr, Xte)
    t1 = time.perf_counter()
    f.get_train_kernel(out=otr); f.get_test_kernel(out=ote)
    t2 = time.perf_counter()
    if rank == 0:
        print(json.dumps({"rep": rep, "ranks": world, "compute_kernel_s": round(t1 - t0, 4), "getters_s": round(t2 - t1, 4), "wall_s": round(t2 - t0, 4)}), flush=True)
    del f
if rank == 0:
    ok, bad = parity_check(X, N_TRAIN, otr, ote, q, parity_sample(N_SEQ, N_TRAIN))
    print(json.dumps({"parity_ok": ok, "cells_differing": bad}))
dist.barrier()
dist.destroy_process_group()
