"""C4 (50 000 x 200 bp, g=16 m=8) resident-input steps under different options: per-class device ms for 384 combinations."""
import itertools, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic, queue_order, N_SEQ, N_TRAIN, SEQ_LEN, G, M
from fastsk_b200 import FastSK, _lib

X = synthetic()
codes = np.ascontiguousarray(X.reshape(-1))
offsets = np.arange(N_SEQ + 1, dtype=np.int64) * SEQ_LEN
q = queue_order()
configs = [dict(seg_dir=1), dict(seg_dir=2), dict(seg_dir=2, dir_blocks=64), dict(seg_dir=2, wave=4), dict(seg_dir=2, wave=1),
           dict(seg_dir=2, count_updates=0), dict(seg_dir=1, count_updates=0)]
if len(sys.argv) > 1:
    configs = [json.loads(a) for a in sys.argv[1:]]
for cfg in configs:
    f = FastSK(G, M, combo_sequence=q, device=0, distributed=False, profile=True)
    for k, v in cfg.items():
        f.set_option(k, v)
    f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), N_TRAIN, N_SEQ - N_TRAIN)
    NC = int(os.environ.get("NCOMB", "384"))
    a = np.ascontiguousarray(q[:NC]); b = np.ascontiguousarray(q[NC:2 * NC])
    f._call("fsk_accumulate_combos", a.ctypes.data_as(_lib.c_i32p), NC, 1)
    s0 = f.stats()
    t0 = time.perf_counter()
    f._call("fsk_accumulate_combos", b.ctypes.data_as(_lib.c_i32p), NC, 1)
    wall = time.perf_counter() - t0
    s1 = f.stats()
    d = {k: round(s1[k] - s0[k], 2) for k in s1 if k.startswith("ms_")}
    print(json.dumps({"cfg": cfg, "wall_ms": round(wall * 1e3, 1), "combos_per_s": round(NC / wall, 1), **d, "seg_mode": s1["seg_mode"], "batch": s1["batch"]}), flush=True)
    del f
