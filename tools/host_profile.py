"""Where the end-to-end time of one compute_kernel goes on the host side (FSK_TRACE=1 prints the library's phases)."""
import os, sys, time
import numpy as np
os.environ["FSK_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic, queue_order, N_TRAIN, G, M, ROOT
from fastsk_b200 import FastSK, FastaUtility
from fastsk_b200.fastsk import pinned_empty, _flatten


def timed(label, fn):
    t0 = time.perf_counter()
    r = fn()
    print(f"[py] {label}: {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr, flush=True)
    return r


which = sys.argv[1] if len(sys.argv) > 1 else "ep300"
if which == "c4":
    X = synthetic()
    q = queue_order()[:int(sys.argv[2]) if len(sys.argv) > 2 else 384]
    for rep in range(2):
        print(f"--- c4 rep {rep}", file=sys.stderr)
        f = timed("FastSK()", lambda: FastSK(G, M, combo_sequence=q, device=0, distributed=False))
        timed("compute_kernel", lambda: f.compute_kernel(X[:N_TRAIN], X[N_TRAIN:]))
        out = pinned_empty((N_TRAIN, N_TRAIN))
        timed("get_train_kernel(pinned out)", lambda: f.get_train_kernel(out=out))
        timed("del", lambda: f.__del__())
else:
    fu = FastaUtility()
    Xtr, _ = fu.read_data(os.path.join(ROOT, "data", "EP300.train.fasta"))
    Xte, _ = fu.read_data(os.path.join(ROOT, "data", "EP300.test.fasta"))
    Atr, Ate = np.array(Xtr, dtype=np.int32), np.array(Xte, dtype=np.int32)
    for rep in range(3):
        print(f"--- ep300 rep {rep} (lists)", file=sys.stderr)
        f = timed("FastSK()", lambda: FastSK(10, 6, seed=0, device=0, distributed=False))
        timed("flatten lists (inside compute_kernel too)", lambda: (_flatten(Xtr), _flatten(Xte)))
        timed("compute_kernel(lists)", lambda: f.compute_kernel(Xtr, Xte))
        timed("get_train_kernel", lambda: f.get_train_kernel())
        timed("get_test_kernel", lambda: f.get_test_kernel())
        print(f"--- ep300 rep {rep} (2-D arrays)", file=sys.stderr)
        f = FastSK(10, 6, seed=0, device=0, distributed=False)
        timed("compute_kernel(arrays)", lambda: f.compute_kernel(Atr, Ate))
        timed("get_train_kernel", lambda: f.get_train_kernel())
