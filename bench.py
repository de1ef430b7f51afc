#!/usr/bin/env python
"""Benchmark of the gapped k-mer kernel-matrix build (BASELINE.json metric: combinations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3], the north-star target --
    X = numpy.random.default_rng(0).integers(1, 5, size=(50000, 200)), train = first 40000, g=16, m=8,
    exact mode, 12 870 combinations in a seed-0 shuffled order.
One step = `--combos-per-step` combinations per GPU run through the whole per-combination path
(pack, sort, segment, accumulate into the resident int64 packed triangle).  Per-GPU work per step is
fixed, so scaling is "weak"; value = combinations processed by all ranks / max-over-ranks device time.

value      inputs resident in HBM, CUDA events on the library's stream, barrier + sync on both sides
e2e        the same number of combinations as one job through the public API with HOST buffers:
           FastSK(...).compute_kernel(Xtrain, Xtest) (H2D of the sequences, sharded build, NCCL all-reduce
           when N > 1, normalisation) + get_train_kernel/get_test_kernel into pinned host memory (D2H)
roofline   dominant kernel class of the step, timed with CUDA events inside the timed region
cpu_baseline / --impl reference
           the reference's own C++ engine (oracle/_ref, compiled from /root/reference in the authoring
           container; else the C port oracle/libfsko.so) on the host cores, on a bounded sample of the
           same synthetic workload (fewer sequences: the reference indexes the triangle with int and needs
           ~6 GB per thread at N = 46 000)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G, M, N_SEQ, N_TRAIN, SEQ_LEN = 16, 8, 50000, 40000, 200
METRIC, UNIT = "gkm_kernel_build_combinations_per_s", "combinations/s"
# dram__bytes_read.sum + dram__bytes_write.sum of accumulate_rows_kernel for one batch of 96 combinations of this workload
# (all rows in one launch, option wave=400), from the ncu --set full capture profiles/r01_ncu_accumulate_rows_batch96_prefetch.txt
# (the L2 prefetch of whole 128-byte lines moves 195 GB where the demand loads alone moved 169 GB)
TRAFFIC_ACC_BATCH96 = 194.906825e9 + 10.006337e9


def synthetic(n=N_SEQ):
    return np.random.default_rng(0).integers(1, 5, size=(N_SEQ, SEQ_LEN), dtype=np.int32)[:n]


def queue_order():
    return np.random.default_rng(0).permutation(comb(G, M)).astype(np.int32)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        try:
            sm = [float(r[0]) for r in rows]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                    "samples": len(rows), "reasons": reasons}
        except Exception as e:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"parse error: {e}"]}


# ------------------------------------------------------------------------------------------------ CPU reference
def host_threads(n_ref):
    """Threads the reference can use: every core, unless its ~ (4 B x pairs + 3 x 4 B x g x windows) per thread
    plus the shared 8 B x pairs would not fit in half of the free RAM."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    pairs = n_ref * (n_ref + 1) // 2
    per_thread = 4 * pairs + 3 * 4 * G * n_ref * (SEQ_LEN - G + 1) + (64 << 20)
    fit = int((avail * 0.5 - 8 * pairs) // per_thread)
    return max(1, min(cores, fit)), cores


def ref_sample_size(budget_s):
    """Sequences in the bounded reference sample so that one combination per thread takes ~budget_s.
    Single-thread cost per combination measured in BASELINE.md: 28.4 s at 46 000, ~quadratic in N."""
    n = int(46000 * min(1.0, (budget_s / 30.0)) ** 0.5)
    return max(2000, min(46000, n // 1000 * 1000))


def reference_step(n_ref, threads, combos):
    """One bounded sample on the host: `threads` reference threads, len(combos) combinations."""
    import oracle
    X = synthetic(n_ref)
    ntr = int(n_ref * 0.8)
    kind = "reference" if oracle.ref_available() else "port"
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)       # the reference prints from its worker threads
    try:
        t0 = time.perf_counter()
        oracle.run("ref" if kind == "reference" else "c", X[:ntr], X[ntr:], G, M, combos, T=threads)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return dt, kind


def cpu_baseline(budget_s):
    n_ref = ref_sample_size(budget_s)
    threads, cores = host_threads(n_ref)
    combos = queue_order()[:threads]
    dt, kind = reference_step(n_ref, threads, combos)
    return {"value": len(combos) / dt, "unit": UNIT, "cores": threads, "host_cores": cores, "kind": kind, "seconds": dt,
            "sample": f"{len(combos)} combinations (one per thread, t={threads}) of the same synthetic set cut to N={n_ref} "
                      f"sequences x {SEQ_LEN} bp, g={G} m={M}, exact mode, incl. g-mer extraction and merge"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = max(4.0, 150.0 / (args.steps + args.warmup))
    n_ref = ref_sample_size(budget)
    threads, cores = host_threads(n_ref)
    q = queue_order()
    kind = "reference"
    for w in range(args.warmup):
        _, kind = reference_step(n_ref, threads, q[w * threads:(w + 1) * threads])
    t_total = 0.0
    for s in range(args.steps):
        o = (args.warmup + s) * threads
        dt, kind = reference_step(n_ref, threads, q[o:o + threads])
        t_total += dt
    value = args.steps * threads / t_total
    sample = (f"per step {threads} combinations (one per thread, t={threads} of {cores} host cores) of the synthetic set cut to "
              f"N={n_ref} sequences x {SEQ_LEN} bp (the reference's int triangle index caps N at 46 341; ~6 GB per thread), "
              f"g={G} m={M}, exact mode")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"synthetic DNA {N_SEQ}x{SEQ_LEN} g={G} m={M} exact (BASELINE configs[3]); reference sample N={n_ref}",
                   "g": G, "m": M, "n_sequences": n_ref, "seq_len": SEQ_LEN, "combos_per_step": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from fastsk_b200 import FastSK, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    X = synthetic()
    Xtr, Xte = X[:N_TRAIN], X[N_TRAIN:]
    queue = queue_order()
    cps = args.combos_per_step
    total_steps = args.warmup + args.steps
    need = total_steps * cps * world
    order = np.resize(queue, need) if need > len(queue) else queue[:need]

    # ---- resident-input arm -------------------------------------------------------------------
    f = FastSK(G, M, combo_sequence=order, device=local, distributed=False, profile=True)
    f.set_option("batch", args.batch)
    f.set_option("acc_path", args.acc_path)
    f.set_option("wave", args.wave)
    codes = np.ascontiguousarray(X.reshape(-1))
    offsets = np.arange(N_SEQ + 1, dtype=np.int64) * SEQ_LEN
    f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), N_TRAIN, N_SEQ - N_TRAIN)
    sp = ctypes.c_void_p()
    f._call("fsk_stream", ctypes.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=f"cuda:{local}")

    def step(i):
        mine = np.ascontiguousarray(order[(i * world + rank) * cps:(i * world + rank + 1) * cps])
        f._call("fsk_accumulate_combos", mine.ctypes.data_as(_lib.c_i32p), len(mine), 0)

    for i in range(args.warmup):
        step(i)
    f._call("fsk_synchronize")
    st0 = f.stats()
    clocks = ClockSampler(local)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for i in range(args.warmup, total_steps):
        step(i)
    ev1.record(stream)
    f._call("fsk_synchronize")
    barrier()
    t_wall1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clock_info = clocks.stop(t_wall0, t_wall1)
    st1 = f.stats()
    value = args.steps * cps * world / (ms * 1e-3)
    d = {k: st1[k] - st0[k] for k in st1 if isinstance(st1[k], (int, float))}
    combos_rank = d["combos_done"]
    updates = d["pair_updates"]
    launches = int(sum_over_ranks(d["kernel_launches"]))

    # finalize (reduction over NVLink + normalisation), timed once, reported beside the rate
    barrier()
    t0 = time.perf_counter()
    if world > 1:
        dist.all_reduce(f.partial_tensor(), op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    f._call("fsk_finalize")
    finalize_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

    peak, peak_src = measured_peak()
    n_pairs, nfeat, rec = st1["n_pairs"], st1["nfeat"], st1["record_bytes"]
    batches = max(1, -(-combos_rank // max(1, st1["batch"])))
    id_bytes = 2 if N_SEQ <= 65536 else 4
    # accumulate (dominant kernel, launched in waves of rows; figures are per batch = one pass over all rows): every unit
    # update streams one sequence id of a run prefix (2 B as u16) and every batch adds each 8-byte cell of the packed
    # triangle once (RED = read + write)
    acc_bytes = float(id_bytes) * updates + 16.0 * n_pairs * batches
    acc_s = d["ms_accumulate"] * 1e-3
    # pack + sort + segment: SURVEY 8(d) formula with the record width actually moved (4 B here, not 8)
    gw_bytes = 4 if G * st1["bits_per_char"] <= 32 else 8
    sort_bytes = combos_rank * nfeat * ((gw_bytes + 4 + rec) + 2 * rec * st1["sort_passes"] + (rec + id_bytes + 8))
    sort_s = (d["ms_pack"] + d["ms_sort"] + d["ms_segment"]) * 1e-3
    roofline = {"kernel": "accumulate_rows_kernel", "bound": "hbm", "achieved": acc_bytes / acc_s / 1e9 if acc_s else None,
                "peak": peak, "unit": "GB/s", "frac": (acc_bytes / acc_s / 1e9 / peak) if acc_s else None, "traffic": (TRAFFIC_ACC_BATCH96 - 16.0 * n_pairs) * st1["batch"] / 96.0 + 16.0 * n_pairs, "traffic_unit": "bytes per batch (ncu dram read + write at batch 96; the id-stream part scaled to this batch, the 16 B x cells flush part not)",
                "achieved_per_launch_bytes": acc_bytes / batches,
                "peak_source": peak_src, "share_of_step": d["ms_accumulate"] / d["ms_total"] if d["ms_total"] else None,
                "algorithmic_bytes": f"{id_bytes} B x unit pair-updates (ids of the run prefixes) + 16 B x packed-triangle cells per batch",
                "launch": "one batch = all row waves of the kernel (launched in waves for L2 locality)",
                "note": "co-limited: ncu (batch 96) shows 205 GB of DRAM traffic for 147 GB algorithmic = 4.7 TB/s (72 % of the measured copy peak; "
                        "whole 128-byte lines of the ~140-byte id prefixes are prefetched into L2 one chunk ahead), the L1/LSU pipe at 86 % and "
                        "shared-memory atomic wavefronts at 67 % (4.2 wavefronts per 32-lane atomic from bank conflicts)",
                "pair_updates_per_s": updates / acc_s if acc_s else None}
    roofline_sort = {"kernel": "pack_hist + onesweep passes + segment", "bound": "hbm", "achieved": sort_bytes / sort_s / 1e9 if sort_s else None,
                     "peak": peak, "unit": "GB/s", "frac": (sort_bytes / sort_s / 1e9 / peak) if sort_s else None,
                     "share_of_step": (d["ms_pack"] + d["ms_sort"] + d["ms_segment"]) / d["ms_total"] if d["ms_total"] else None,
                     "algorithmic_bytes": f"per combination and window: ({gw_bytes}+4+{rec}) pack + 2 x {rec} x {st1['sort_passes']} passes + "
                                          f"({rec}+{id_bytes}+8) segment",
                     "per_stage_gbs": {
                         "pack": combos_rank * nfeat * (gw_bytes + 4 + rec) / (d["ms_pack"] * 1e-3) / 1e9 if d["ms_pack"] else None,
                         "sort": combos_rank * nfeat * 2 * rec * st1["sort_passes"] / (d["ms_sort"] * 1e-3) / 1e9 if d["ms_sort"] else None,
                         "segment": combos_rank * nfeat * (rec + id_bytes + 8) / (d["ms_segment"] * 1e-3) / 1e9 if d["ms_segment"] else None}}
    phase_ms = {k[3:]: d[k] / args.steps for k in d if k.startswith("ms_")}
    del f
    torch.cuda.empty_cache()

    # ---- end-to-end arm: public API, host buffers ------------------------------------------------
    job = order[args.warmup * cps * world:]
    pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    out_tr = pin(N_TRAIN * N_TRAIN).reshape(N_TRAIN, N_TRAIN) if rank == 0 else None
    out_te = pin((N_SEQ - N_TRAIN) * N_TRAIN).reshape(N_SEQ - N_TRAIN, N_TRAIN) if rank == 0 else None
    Xtr_p = torch.from_numpy(Xtr.copy()).pin_memory().numpy()
    Xte_p = torch.from_numpy(Xte.copy()).pin_memory().numpy()

    def e2e_job(q):
        t = [time.perf_counter()]
        fe = FastSK(G, M, combo_sequence=q, device=local)
        fe.set_option("batch", args.batch)
        fe.set_option("acc_path", args.acc_path)
        fe.set_option("wave", args.wave)
        fe.compute_kernel(Xtr_p, Xte_p)
        t.append(time.perf_counter())
        if rank == 0:
            fe.get_train_kernel(out=out_tr)
            fe.get_test_kernel(out=out_te)
        t.append(time.perf_counter())
        return {"compute_kernel_s": t[1] - t[0], "get_kernels_d2h_s": t[2] - t[1]}, fe

    _, fe = e2e_job(job[:world * min(cps, 8)])        # warm-up: allocator, NCCL channels
    del fe
    barrier()
    t0 = time.perf_counter()
    e2e_parts, fe = e2e_job(job)                      # both kernels are in host memory when this returns
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    del fe                                            # teardown (cudaFree of ~40 GB) is not part of the job's result
    barrier()
    e2e = {"value": len(job) / e2e_s, "unit": UNIT, "seconds": e2e_s, "combinations": int(len(job)), "parts_rank0": e2e_parts,
           "h2d_bytes_per_step": int(X.nbytes // args.steps), "d2h_bytes_per_step": int((out_tr.nbytes + out_te.nbytes) // args.steps) if rank == 0 else 0,
           "what": "FastSK(g,m,combo_sequence=<the timed region's combinations>).compute_kernel(Xtrain, Xtest) from pinned host int32 + "
                   "get_train_kernel/get_test_kernel into pinned host fp64 (one job = all steps; bytes are per-step shares)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline(args.cpu_budget) if (world == 1 and not args.no_cpu_baseline) else None
    other = other_workloads(local) if world == 1 else None
    if other is not None and not args.no_skewed:
        other["skewed"] = skewed_workload(local)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": f"synthetic DNA {N_SEQ}x{SEQ_LEN} g={G} m={M} exact (BASELINE configs[3]: 12870 combinations)",
                   "n_train": N_TRAIN, "n_test": N_SEQ - N_TRAIN, "g": G, "m": M, "combos_per_step_per_gpu": cps, "batch": st1["batch"],
                   "parallelism": f"combinations sharded over {world} GPU(s), one final NCCL all-reduce",
                   "l2": "inputs larger than L2: each step streams > 10 GB (packed int64 triangle + per-slot records)",
                   "record_bytes": rec, "key_bits": st1["key_bits"], "sort_passes": st1["sort_passes"]},
        "roofline": roofline, "roofline_sort": roofline_sort, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clock_info, "phase_ms_per_step": phase_ms, "pair_updates_per_s": updates * world / (ms * 1e-3),
        "finalize_ms": finalize_ms, "projected_full_build_s": comb(G, M) / value + finalize_ms * 1e-3,
        "other_workloads": other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_workloads(device):
    """BASELINE configs[1] beside the headline: the bundled EP300 TFBS set (4000 x 100 bp DNA, g=10 m=6, exact, all 210
    combinations) through the public API, once per accumulate path.  With 256 distinct k-mers per combination the
    update is a dense contraction: acc_path 3 builds K += C C^T with tcgen05 MMAs (fsk_dense.cuh); acc_path 2 is the
    sort + shared-memory row path of the headline workload.  Device milliseconds are CUDA-event spans of the library."""
    tr, te = os.path.join(ROOT, "data", "EP300.train.fasta"), os.path.join(ROOT, "data", "EP300.test.fasta")
    if not (os.path.exists(tr) and os.path.exists(te)):
        return None
    from fastsk_b200 import FastSK, FastaUtility
    fu = FastaUtility()
    Xtr, _ = fu.read_data(tr)
    Xte, _ = fu.read_data(te)
    out = {"workload": "EP300 TFBS DNA train+test (BASELINE configs[1]): 4000 sequences x 100 bp, g=10 m=6, exact, 210 combinations"}
    for path, tag in ((2, "rows"), (3, "dense_tensor_core")):
        best = None
        for _ in range(3):
            f = FastSK(10, 6, seed=0, device=device, distributed=False, profile=True)
            f.set_option("acc_path", path)
            t0 = time.perf_counter()
            f.compute_kernel(Xtr, Xte)
            Ktr = f.get_train_kernel()
            wall = time.perf_counter() - t0
            st = f.stats()
            dev_ms = st["ms_total"]
            row = {"e2e_s": wall, "device_ms": dev_ms, "combinations_per_s_device": st["combos_done"] / (dev_ms * 1e-3),
                   "combinations_per_s_e2e": st["combos_done"] / wall, "ms_accumulate": st["ms_accumulate"], "ms_pack": st["ms_pack"],
                   "kernel_launches": st["kernel_launches"], "trace": float(np.trace(Ktr))}
            if path == 3 and st["ms_accumulate"]:
                n, kdim = st["n_seq"], 256 * st["combos_done"]
                row["tensor_tflops"] = 2.0 * (n * (n + 128) / 2.0) * kdim / (st["ms_accumulate"] * 1e-3) / 1e12
            if best is None or row["device_ms"] < best["device_ms"]:
                best = row
            del f
        out[tag] = best
    return out


def skewed_dna(n, L, seed=1, motif_len=12, frac=0.5):
    """SURVEY 8(d) skewed variant of the headline set: first-order Markov, GC-rich, one planted 12-mer in half of the
    sequences -- heavy-tailed run lengths where the uniform set has ~141 records in every run."""
    rng = np.random.default_rng(seed)
    P = np.array([[0.10, 0.40, 0.40, 0.10], [0.05, 0.45, 0.45, 0.05], [0.05, 0.45, 0.45, 0.05], [0.10, 0.40, 0.40, 0.10]])
    cdf = np.cumsum(P, axis=1)
    X = np.empty((n, L), dtype=np.int32)
    X[:, 0] = rng.integers(0, 4, size=n)
    u = rng.random((n, L))
    for t in range(1, L):
        X[:, t] = (u[:, t, None] > cdf[X[:, t - 1]]).sum(axis=1)
    motif = rng.integers(0, 4, size=motif_len)
    for i in np.flatnonzero(rng.random(n) < frac):
        p = int(rng.integers(0, L - motif_len + 1))
        X[i, p:p + motif_len] = motif
    return (np.minimum(X, 3) + 1).astype(np.int32)


def skewed_workload(device, combos=96):
    """The headline shape (50000 x 200 bp, g=16 m=8) on the skewed set, one batch of combinations with inputs resident:
    the row path alone against the row path with its heavy runs (> 0.05 N records) contracted on the tensor cores."""
    from fastsk_b200 import FastSK, _lib
    X = skewed_dna(N_SEQ, SEQ_LEN)
    order = queue_order()[:2 * combos]
    codes = np.ascontiguousarray(X.reshape(-1))
    offsets = np.arange(N_SEQ + 1, dtype=np.int64) * SEQ_LEN
    out = {"workload": f"skewed synthetic DNA {N_SEQ}x{SEQ_LEN} (Markov GC-rich + planted 12-mer), g={G} m={M}, {combos} combinations timed after {combos} warm-up"}
    for tau, tag in ((-1, "rows_only"), (0, "rows_plus_heavy_runs_on_tensor_cores")):
        f = FastSK(G, M, combo_sequence=order, device=device, distributed=False, profile=True)
        f.set_option("heavy_tau", tau)
        f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), N_TRAIN, N_SEQ - N_TRAIN)
        f._call("fsk_accumulate_combos", order[:combos].ctypes.data_as(_lib.c_i32p), combos, 1)
        s0 = f.stats()
        q = np.ascontiguousarray(order[combos:])
        f._call("fsk_accumulate_combos", q.ctypes.data_as(_lib.c_i32p), combos, 1)
        s1 = f.stats()
        ms = s1["ms_total"] - s0["ms_total"]
        out[tag] = {"device_ms": ms, "combinations_per_s": combos / (ms * 1e-3), "heavy_runs": s1["heavy_runs"] - s0["heavy_runs"],
                    "heavy_tau": s1["heavy_tau"], "pair_updates": s1["pair_updates"] - s0["pair_updates"]}
        del f
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--combos-per-step", type=int, default=384)
    ap.add_argument("--batch", type=int, default=0, help="combinations per launch group (0 = auto)")
    ap.add_argument("--acc-path", type=int, default=0, help="0 auto, 1 global RED, 2 shared-memory rows, 3 dense tensor-core")
    ap.add_argument("--wave", type=int, default=32, help="accumulate launch = wave x resident CTAs rows")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of reference CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-skewed", action="store_true", help="skip the skewed-set section of other_workloads (~6 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
