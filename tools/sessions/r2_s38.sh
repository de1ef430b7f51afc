#!/bin/bash
# r2 session 38 (1 GPU): byte operands / int32 accumulators for the register-form Welford contraction -- variance tests, EP300 timing both forms, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "approx or variance or speculated or registers or golden or column_windows" > gpurun_out/r2s38_pytest.txt 2>&1
tail -4 gpurun_out/r2s38_pytest.txt
timeout 300 python - > gpurun_out/r2s38_ep300_approx.txt 2>&1 <<'PY'
import json, sys, time
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for regs, u8 in ((1, 1), (1, 0)):
    for kw in (dict(t=1, max_iters=50), dict(t=20, max_iters=50), dict(t=1, max_iters=210)):
        best = None
        for rep in range(4):
            f = FastSK(10, 6, approx=True, seed=0, device=0, distributed=False, profile=True, **kw)
            f.set_option("wf_regs", regs)
            f.set_option("wf_u8", u8)
            t0 = time.perf_counter(); f.compute_kernel(Xtr, Xte); wall = time.perf_counter() - t0
            st = f.stats()
            row = {"wf_regs": regs, "wf_u8": u8, "cfg": kw, "wall_ms": round(wall * 1e3, 2), "device_ms": round(st["ms_total"], 3), "combos": st["combos_done"],
                   "combos_per_s_device": round(st["combos_done"] / (st["ms_total"] * 1e-3)), "launches": st["kernel_launches"], "stdevs": len(f.get_stdevs()),
                   "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")}}
            if best is None or row["device_ms"] < best["device_ms"]: best = row
        print(json.dumps(best), flush=True)
PY
cat gpurun_out/r2s38_ep300_approx.txt
cat > /tmp/ep300_approx.py <<'PY'
import sys
sys.path.insert(0, ".")
from fastsk_b200 import FastSK, FastaUtility
fu = FastaUtility()
Xtr, _ = fu.read_data("data/EP300.train.fasta"); Xte, _ = fu.read_data("data/EP300.test.fasta")
for rep in range(2):
    f = FastSK(10, 6, t=1, approx=True, max_iters=50, seed=0, device=0, distributed=False)
    f.compute_kernel(Xtr, Xte)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:syrk_tc_welford -c 1 -f -o gpurun_out/r2s38_welford python /tmp/ep300_approx.py > gpurun_out/r2s38_ncu.log 2>&1
tail -2 gpurun_out/r2s38_ncu.log
timeout 200 python bench.py --steps 1 --warmup 3 --combos 384 --no-parity --no-cpu-baseline --no-skewed > gpurun_out/r2s38_bench_short.json 2> gpurun_out/r2s38_bench_short.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s38_bench_short.json").read().strip().splitlines()[-1])
    print("value", d["value"], {k: d["other_workloads"][k] for k in d.get("other_workloads", {}) if "approx" in k})
except Exception as e:
    print("bench:", e)
PY
