"""The full configs[3] build from ONE plain process on every visible GPU (fsk_set_devices): wall seconds, parity sample."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic, queue_order, parity_sample, parity_check, N_SEQ, N_TRAIN, G, M
from fastsk_b200 import FastSK
from fastsk_b200.fastsk import pinned_empty

X = synthetic()
q = queue_order() if len(sys.argv) < 2 else queue_order()[:int(sys.argv[1])]
n_test = N_SEQ - N_TRAIN
Xtr, Xte = pinned_empty((N_TRAIN, 200), np.int32), pinned_empty((n_test, 200), np.int32)
Xtr[:], Xte[:] = X[:N_TRAIN], X[N_TRAIN:]
otr, ote = pinned_empty((N_TRAIN, N_TRAIN)), pinned_empty((n_test, N_TRAIN))
for rep in range(3):
    t0 = time.perf_counter()
    f = FastSK(G, M, combo_sequence=q, devices="all")
    f.compute_kernel(Xtr, Xte)
    t1 = time.perf_counter()
    f.get_train_kernel(out=otr)
    f.get_test_kernel(out=ote)
    t2 = time.perf_counter()
    st = f.stats()
    print(json.dumps({"rep": rep, "devices": st["n_devices"], "combinations": int(st["combos_done"]), "compute_kernel_s": round(t1 - t0, 4),
                      "getters_s": round(t2 - t1, 4), "wall_s": round(t2 - t0, 4), "combinations_per_s": round(len(q) / (t2 - t0), 1)}), flush=True)
    del f
ok, bad = parity_check(X, N_TRAIN, otr, ote, q, parity_sample(N_SEQ, N_TRAIN))
print(json.dumps({"parity_ok": ok, "cells_differing": bad}))
