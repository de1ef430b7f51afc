"""The skewed configs[3]-shaped set (bench.skewed_dna): per-class device ms and counters for 96 combinations."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import skewed_dna, queue_order, N_SEQ, N_TRAIN, SEQ_LEN, G, M
from fastsk_b200 import FastSK, _lib
X = skewed_dna(N_SEQ, SEQ_LEN)
codes = np.ascontiguousarray(X.reshape(-1)); offsets = np.arange(N_SEQ + 1, dtype=np.int64) * SEQ_LEN
q = queue_order()
for cfg in [json.loads(a) for a in sys.argv[1:]] or [{}]:
    f = FastSK(G, M, combo_sequence=q, device=0, distributed=False, profile=True)
    for k, v in cfg.items():
        f.set_option(k, v)
    f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), N_TRAIN, N_SEQ - N_TRAIN)
    a = np.ascontiguousarray(q[:96]); b = np.ascontiguousarray(q[96:192])
    f._call("fsk_accumulate_combos", a.ctypes.data_as(_lib.c_i32p), 96, 1)
    s0 = f.stats()
    f._call("fsk_accumulate_combos", b.ctypes.data_as(_lib.c_i32p), 96, 1)
    s1 = f.stats()
    d = {k: round(s1[k] - s0[k], 2) for k in s1 if k.startswith("ms_")}
    print(json.dumps({"cfg": cfg, **d, "entries": s1["entries"] - s0["entries"], "runs": s1["runs"] - s0["runs"],
                      "pair_updates_rows": s1["pair_updates"] - s0["pair_updates"], "heavy_runs": s1["heavy_runs"] - s0["heavy_runs"],
                      "heavy_tau": s1["heavy_tau"], "records": 96 * s1["nfeat"]}), flush=True)
    del f
