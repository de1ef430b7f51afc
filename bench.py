#!/usr/bin/env python
"""Benchmark of the gapped k-mer kernel-matrix build (BASELINE.json metric: combinations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3], the north-star target --
    X = numpy.random.default_rng(0).integers(1, 5, size=(50000, 200)), train = first 40000, g=16, m=8,
    exact mode, 12 870 combinations in a seed-0 shuffled order.
One step = ONE COMPLETE BUILD of that kernel: all 12 870 combinations, dealt round-robin to the N ranks (pack, sort,
segment, accumulate into each rank's int64 packed triangle), then the normalisation of each rank's share of the
output rows, which merges the partial kernels of all ranks over NVLink as it reads them.  Total work per step is
fixed, so scaling is "strong"; value = 12 870 x steps / max-over-ranks device time, wall_s_per_build beside it.

value      inputs resident in HBM, CUDA events on the library's stream, barrier + sync on both sides
e2e        the same build as one job through the public API with HOST buffers: FastSK(...).compute_kernel(Xtrain,
           Xtest) (H2D of the sequences on every rank, sharded build, peer merge + normalisation) +
           get_train_kernel/get_test_kernel into pinned host memory every rank maps (parallel D2H); wall_s
parity     after the timed region the job's own output is compared, bit for bit, with the C oracle on a sample of
           sequences (K_ij depends on sequences i and j only): parity_ok
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
# dram__bytes_read.sum + dram__bytes_write.sum of accumulate_rows_kernel for one batch of 48 combinations of this workload
# (all rows in one launch, option wave=400), from the ncu --set full capture profiles/r02_ncu_accumulate_final.txt
TRAFFIC_ACC_BATCH48 = 97.352445e9 + 10.004478e9


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


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return 2250.0, "nominal (B200_PROFILING.md)"


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


def reference_step(n_ref, threads, combos, keep=False):
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
        K, _, _ = oracle.run("ref" if kind == "reference" else "c", X[:ntr], X[ntr:], G, M, combos, T=threads)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return (dt, kind, K) if keep else (dt, kind)


def cpu_baseline(budget_s, device):
    """The reference on the host cores on a bounded sample, and -- same sequences, same combinations -- this library:
    the full unnormalised matrix must be equal, and the two times are a like-for-like ratio."""
    n_ref = ref_sample_size(budget_s)
    threads, cores = host_threads(n_ref)
    combos = queue_order()[:threads]
    dt, kind, K_ref = reference_step(n_ref, threads, combos, keep=True)
    cpu = {"value": len(combos) / dt, "unit": UNIT, "cores": threads, "host_cores": cores, "kind": kind, "seconds": dt,
           "sample": f"{len(combos)} combinations (one per thread, t={threads}) of the same synthetic set cut to N={n_ref} "
                     f"sequences x {SEQ_LEN} bp, g={G} m={M}, exact mode, incl. g-mer extraction and merge"}
    from fastsk_b200 import FastSK
    X = synthetic(n_ref)
    ntr = int(n_ref * 0.8)
    best = None
    for _ in range(2):                     # the second run re-uses the device blocks of the first
        f = FastSK(G, M, combo_sequence=combos, device=device, distributed=False, profile=True)
        t0 = time.perf_counter()
        f.compute_kernel(X[:ntr], X[ntr:])
        t1 = time.perf_counter() - t0
        best = t1 if best is None else min(best, t1)
        dev_ms = f.stats()["ms_total"]
        K_gpu = f.get_unnormalised(np.float64)
        del f
    same = {"workload": cpu["sample"], "parity_ok_full_matrix": bool(np.array_equal(K_gpu, K_ref)), "cells": int(K_ref.size),
            "reference_s": dt, "b200_compute_kernel_s": best, "b200_device_ms": dev_ms, "ratio_e2e": dt / best,
            "ratio_device": dt / (dev_ms * 1e-3) if dev_ms else None,
            "what": "same sequences, same combinations, same mode on both sides; b200 time = compute_kernel from host arrays "
                    "(upload + build + normalisation), so 16 combinations mostly measure its fixed costs"}
    return cpu, same


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
def parity_sample(n_seq, n_train, k=48, seed=5):
    """Sequences whose K sub-block is checked against the oracle after the timed region: row 0, the last train row, the
    first test row, row N-1 and random others from both blocks."""
    rng = np.random.default_rng(seed)
    fixed = [0, n_train - 1, n_train, n_seq - 1]
    rest = rng.choice(np.setdiff1d(np.arange(n_seq), fixed), size=k - len(fixed), replace=False)
    return np.sort(np.concatenate([fixed, rest])).astype(np.int64)


def parity_check(X, n_train, train, test, queue, sample):
    """K_ij depends on sequences i and j only, so the oracle run on the sampled sequences alone gives the cells of the
    full build at those rows and columns: the normalised values must be bit-equal (exact mode: integer sums, then the
    same IEEE mul / sqrt / div)."""
    import oracle
    sub = X[sample]
    K, _, _ = oracle.run("c", sub.tolist(), [], G, M, queue, T=1, normalise=True)
    S = oracle.unpack(K, len(sample))
    cols = sample[sample < n_train]
    got = np.empty((len(sample), len(cols)))
    for a, i in enumerate(sample):
        got[a] = train[i, cols] if i < n_train else test[i - n_train, cols]
    want = S[:, :len(cols)]
    return bool(np.array_equal(got, want)), int((got != want).sum())


def run_b200(args):
    import torch
    import torch.distributed as dist
    from fastsk_b200 import FastSK, _lib
    from fastsk_b200.fastsk import pinned_empty, shared_output

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

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX)   # noqa: E731
    sum_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.SUM)   # noqa: E731

    X = synthetic()
    n_test = N_SEQ - N_TRAIN
    queue = queue_order() if args.combos == 0 else queue_order()[:args.combos]
    n_combos = len(queue)

    # ---- resident-input arm: one step = one complete exact build -----------------------------------
    # reset -> this rank's shard of the combinations (pack, sort, segment, accumulate) -> barrier -> normalisation of this
    # rank's output rows, which sums the partial kernels of all ranks over NVLink as it reads them -> barrier
    f = FastSK(G, M, combo_sequence=queue, device=local, distributed=False, profile=True)
    f.set_option("batch", args.batch)
    f.set_option("acc_path", args.acc_path)
    f.set_option("wave", args.wave)
    f.set_shard(rank, world)
    codes = np.ascontiguousarray(X.reshape(-1))
    offsets = np.arange(N_SEQ + 1, dtype=np.int64) * SEQ_LEN
    f._call("fsk_upload", codes.ctypes.data_as(_lib.c_i32p), offsets.ctypes.data_as(_lib.c_i64p), N_TRAIN, n_test)
    merge = "none (one GPU)"
    if world > 1:
        if f._exchange_peers(dist, world):
            merge = "peer loads over NVLink inside the normalisation kernel (CUDA IPC)"
        else:
            merge = "NCCL all-reduce (CUDA IPC unavailable)"
    sp = ctypes.c_void_p()
    f._call("fsk_stream", ctypes.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=f"cuda:{local}")
    t_build = t_final = 0.0

    def step(timed):
        nonlocal t_build, t_final
        t0 = time.perf_counter()
        f._call("fsk_build_partial")
        if world > 1:
            dist.barrier()                       # every rank's partial kernel is complete (the slowest shard sets the time)
        t1 = time.perf_counter()
        if world > 1:
            if merge.startswith("NCCL"):
                dist.all_reduce(f.partial_tensor(), op=dist.ReduceOp.SUM)
                torch.cuda.synchronize()
        f._call("fsk_finalize")
        if world > 1:
            dist.barrier()
        if timed:
            t_build += t1 - t0
            t_final += time.perf_counter() - t1

    # pair updates of one build (deterministic): counted during the first warm-up build only, the timed builds run without
    # the counters
    sw = f.stats()
    for w in range(args.warmup):
        step(False)
        if w == 0:
            updates_per_build = f.stats()["pair_updates"] - sw["pair_updates"]
            f.set_option("count_updates", 0)
    f._call("fsk_synchronize")
    st0 = f.stats()
    clocks = ClockSampler(local)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step(True)
    ev1.record(stream)
    f._call("fsk_synchronize")
    barrier()
    t_wall1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clock_info = clocks.stop(t_wall0, t_wall1)
    st1 = f.stats()
    value = args.steps * n_combos / (ms * 1e-3)
    d = {k: st1[k] - st0[k] for k in st1 if isinstance(st1[k], (int, float))}
    combos_rank = d["combos_done"]
    updates = updates_per_build * args.steps
    launches = int(sum_over_ranks(d["kernel_launches"]))
    updates_all = sum_over_ranks(float(updates))
    build_ms = max_over_ranks(1e3 * t_build / args.steps)
    merge_norm_ms = max_over_ranks(1e3 * t_final / args.steps)

    peak, peak_src = measured_peak()
    n_pairs, nfeat, rec = st1["n_pairs"], st1["nfeat"], st1["record_bytes"]
    batches = max(1, -(-combos_rank // max(1, st1["batch"])))
    id_bytes = 2 if N_SEQ <= 65000 else 4
    # accumulate (dominant kernel; figures per batch = one pass over all rows): every unit update streams one sequence id
    # of a run prefix (2 B as u16) and every batch adds each 8-byte cell of the packed triangle once (RED = read + write)
    acc_bytes = float(id_bytes) * updates + 16.0 * n_pairs * batches
    acc_s = d["ms_accumulate"] * 1e-3
    traffic_batch = (TRAFFIC_ACC_BATCH48 - 16.0 * n_pairs) * st1["batch"] / 48.0 + 16.0 * n_pairs
    gw_bytes = 4 if G * st1["bits_per_char"] <= 32 else 8
    # pre-pass bytes per window and combination.  pack: the g-mer word and window -> sequence table are shared by all
    # slots of a batch and come from L2 (ncu: 0.07 GB of DRAM reads per batch), so its HBM traffic is the record it writes;
    # sort: one read + one write per pass; segment: read the sorted record, write the id (+ the task of the task-list form)
    seg_bytes = rec + id_bytes + (8 if (st1.get("seg_mode", 0) & 3) != 1 else 0)      # (the directory form files no task per record)
    stage_bytes = {"pack": rec, "sort": 2 * rec * st1["sort_passes"], "segment": seg_bytes}
    sort_bytes = combos_rank * nfeat * sum(stage_bytes.values())
    sort_s = (d["ms_pack"] + d["ms_sort"] + d["ms_segment"]) * 1e-3
    roofline = {"kernel": "accumulate_rows_kernel", "bound": "hbm", "achieved": acc_bytes / acc_s / 1e9 if acc_s else None,
                "peak": peak, "unit": "GB/s", "frac": (acc_bytes / acc_s / 1e9 / peak) if acc_s else None,
                "traffic": traffic_batch,
                "traffic_unit": "bytes per batch (ncu dram read + write at batch 48; the id-stream part scaled to this batch, the 16 B x cells flush part not)",
                "frac_measured_traffic": (traffic_batch * batches / acc_s / 1e9 / peak) if acc_s else None,
                "achieved_per_launch_bytes": acc_bytes / batches,
                "peak_source": peak_src, "share_of_step": d["ms_accumulate"] / d["ms_total"] if d["ms_total"] else None,
                "algorithmic_bytes": f"{id_bytes} B x unit pair-updates (ids of the run prefixes) + 16 B x packed-triangle cells per batch",
                "launch": "one batch = all row waves of the kernel (launched in waves for L2 locality)",
                "note": "co-limited: shared-memory atomic wavefronts (4.2 per 32-lane atomic from bank conflicts), the L1/LSU pipe and DRAM "
                        "(whole 128-byte lines of the ~140-byte id prefixes); see profiles/ for the ncu captures"}
    roofline_sort = {"kernel": "pack_hist + onesweep passes + segment", "bound": "hbm", "achieved": sort_bytes / sort_s / 1e9 if sort_s else None,
                     "peak": peak, "unit": "GB/s", "frac": (sort_bytes / sort_s / 1e9 / peak) if sort_s else None,
                     "frac_survey_8d_nominal": (combos_rank * (N_SEQ * SEQ_LEN + nfeat * 8.0 * (3 + 2 * st1["sort_passes"])) / sort_s / 1e9 / peak) if sort_s else None,
                     "share_of_step": (d["ms_pack"] + d["ms_sort"] + d["ms_segment"]) / d["ms_total"] if d["ms_total"] else None,
                     "algorithmic_bytes": "per combination and window: " + " + ".join(f"{v} {k}" for k, v in stage_bytes.items()) +
                                          " (HBM bytes: pack's inputs are L2 hits shared by the slots of a batch); frac_survey_8d_nominal uses "
                                          "SURVEY 8(d)'s B_combo with 8-byte records",
                     "per_stage_gbs": {k: (combos_rank * nfeat * v / (d["ms_" + k] * 1e-3) / 1e9 if d["ms_" + k] else None)
                                       for k, v in stage_bytes.items()}}
    phase_ms = {k[3:]: d[k] / args.steps for k in d if k.startswith("ms_")}
    f._call("fsk_release_peers")
    barrier()
    del f
    torch.cuda.empty_cache()

    # ---- end-to-end arm: the same build as ONE job through the public API, host buffers in and out -------------------
    if world > 1:
        out_tr, out_te = shared_output(N_TRAIN, N_TRAIN, dist), shared_output(n_test, N_TRAIN, dist)
    else:
        out_tr, out_te = pinned_empty((N_TRAIN, N_TRAIN)), pinned_empty((n_test, N_TRAIN))
    Xtr_p, Xte_p = pinned_empty((N_TRAIN, SEQ_LEN), np.int32), pinned_empty((n_test, SEQ_LEN), np.int32)
    Xtr_p[:], Xte_p[:] = X[:N_TRAIN], X[N_TRAIN:]

    def e2e_job(q):
        t = [time.perf_counter()]
        fe = FastSK(G, M, combo_sequence=q, device=local)
        fe.set_option("batch", args.batch)
        fe.set_option("acc_path", args.acc_path)
        fe.set_option("wave", args.wave)
        fe.compute_kernel(Xtr_p, Xte_p)
        t.append(time.perf_counter())
        fe.get_train_kernel(out=out_tr)
        fe.get_test_kernel(out=out_te)
        t.append(time.perf_counter())
        return {"compute_kernel_s": t[1] - t[0], "get_kernels_d2h_s": t[2] - t[1]}, fe

    _, fe = e2e_job(queue[:world * 8])               # warm-up: device blocks, NCCL channels, IPC
    del fe
    e2e_runs = []
    for _ in range(1 if world == 1 else 2):
        barrier()
        t0 = time.perf_counter()
        e2e_parts, fe = e2e_job(queue)               # both kernels are in host memory when this returns
        torch.cuda.synchronize()
        e2e_runs.append(max_over_ranks(time.perf_counter() - t0))
        del fe
    e2e_s = min(e2e_runs)
    barrier()
    d2h = int(sum_over_ranks(float(out_tr.nbytes // world + out_te.nbytes // world)))
    e2e = {"value": n_combos / e2e_s, "unit": UNIT, "wall_s": e2e_s, "runs_s": e2e_runs, "combinations": int(n_combos),
           "parts_rank0": e2e_parts, "h2d_bytes_per_step": int(X.nbytes) * world, "d2h_bytes_per_step": d2h,
           "what": "one job = one step: FastSK(g, m, combo_sequence=<all combinations>).compute_kernel(Xtrain, Xtest) from pinned host "
                   "int32 (every rank uploads the sequences) + get_train_kernel / get_test_kernel into host fp64 (every rank copies its "
                   "rows over its own PCIe link into one buffer all ranks map)"}

    # ---- parity inside the run: the job's own output against the oracle ------------------------------------------------
    parity = None
    if rank == 0 and not args.no_parity:
        sample = parity_sample(N_SEQ, N_TRAIN)
        t0 = time.perf_counter()
        ok, bad = parity_check(X, N_TRAIN, out_tr, out_te, queue, sample)
        parity = {"parity_ok": ok, "cells": int(len(sample) * (sample < N_TRAIN).sum()), "cells_differing": bad,
                  "what": f"normalised kernel of the e2e job at {len(sample)} sampled sequences (rows 0, n_train-1, n_train, N-1 and "
                          f"random others) against the C oracle run on those sequences, bit-equal",
                  "oracle_s": time.perf_counter() - t0}
    if rank != 0:
        if world > 1:
            aimed_workload(local, dist)
            dist.destroy_process_group()
        return
    cpu = same = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, same = cpu_baseline(args.cpu_budget, local)
    other = other_workloads(local) if world == 1 else {}
    if other is not None:
        if world == 1 and not args.no_skewed:
            other["skewed"] = skewed_workload(local)
        other["same_config"] = same
        other["aimed_approx"] = aimed_workload(local, dist if world > 1 else None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "wall_s_per_build": ms * 1e-3 / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"synthetic DNA {N_SEQ}x{SEQ_LEN} g={G} m={M} exact, the full build (BASELINE configs[3]: {n_combos} combinations per step)",
                   "n_train": N_TRAIN, "n_test": n_test, "g": G, "m": M, "combinations_per_step": n_combos, "batch": st1["batch"],
                   "parallelism": f"combinations dealt round-robin to {world} GPU(s); merge: {merge}; every rank normalises 1/{world} of the rows",
                   "l2": "inputs larger than L2: each batch streams > 10 GB (packed int64 triangle + per-slot records)",
                   "record_bytes": rec, "key_bits": st1["key_bits"], "sort_passes": st1["sort_passes"]},
        "roofline": roofline, "roofline_sort": roofline_sort, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
        "parity_ok": None if parity is None else parity["parity_ok"], "gpu_launches": launches,
        "clocks": clock_info, "phase_ms_per_step": phase_ms, "pair_updates_per_s": updates_all / (ms * 1e-3),
        "step_parts_ms": {"build_shard": build_ms, "merge_and_normalise": merge_norm_ms,
                          "what": "host-timed parts of one step, max over ranks: the shard's pack/sort/segment/accumulate up to the barrier "
                                  "that ends the slowest shard; then the normalisation of this rank's rows with the merge of all ranks' partial "
                                  "kernels fused into its loads (peer reads over NVLink; the limiting exchange of the step) + closing barrier"},
        "other_workloads": other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_workloads(device):
    """BASELINE configs[1] beside the headline: the bundled EP300 TFBS set (4000 x 100 bp DNA, g=10 m=6, exact, all 210
    combinations) through the public API, once per accumulate path.  With 256 distinct k-mers per combination the
    update is a dense contraction: acc_path 3 builds K += C C^T with tcgen05 MMAs (fsk_dense.cuh; byte operands); acc_path 2 is the
    sort + shared-memory row path of the headline workload.  Device milliseconds are CUDA-event spans of the library."""
    tr, te = os.path.join(ROOT, "data", "EP300.train.fasta"), os.path.join(ROOT, "data", "EP300.test.fasta")
    if not (os.path.exists(tr) and os.path.exists(te)):
        return None
    from fastsk_b200 import FastSK, FastaUtility
    from fastsk_b200.fastsk import pinned_empty
    fu = FastaUtility()
    ctr, otr, _ = fu.read_encoded(tr)              # flat codes + offsets: the ingest path that skips Python lists
    cte, ote, _ = fu.read_encoded(te)
    Xtr, Xte = (ctr, otr), (cte, ote)
    ntr, nte = len(otr) - 1, len(ote) - 1
    out_tr, out_te = pinned_empty((ntr, ntr)), pinned_empty((nte, ntr))
    out = {"workload": "EP300 TFBS DNA train+test (BASELINE configs[1]): 4000 sequences x 100 bp, g=10 m=6, exact, 210 combinations; "
                       "e2e = compute_kernel from flat arrays (FastaUtility.read_encoded) + both kernels into pinned host arrays"}
    for path, tag in ((2, "rows"), (3, "dense_tensor_core")):
        best = None
        for _ in range(4):
            f = FastSK(10, 6, seed=0, device=device, distributed=False, profile=True)
            f.set_option("acc_path", path)
            t0 = time.perf_counter()
            f.compute_kernel(Xtr, Xte)
            Ktr = f.get_train_kernel(out=out_tr)
            f.get_test_kernel(out=out_te)
            wall = time.perf_counter() - t0
            st = f.stats()
            dev_ms = st["ms_total"]
            row = {"e2e_s": wall, "device_ms": dev_ms, "combinations_per_s_device": st["combos_done"] / (dev_ms * 1e-3),
                   "combinations_per_s_e2e": st["combos_done"] / wall, "ms_accumulate": st["ms_accumulate"], "ms_pack": st["ms_pack"],
                   "kernel_launches": st["kernel_launches"], "trace": float(np.trace(Ktr))}
            if path == 3 and st["ms_accumulate"]:
                n, kdim = st["n_seq"], 256 * st["combos_done"]
                row["tensor_tflops"] = 2.0 * (n * (n + 128) / 2.0) * kdim / (st["ms_accumulate"] * 1e-3) / 1e12
                row["tensor_operands"] = "u8 x u8 -> s32 (tcgen05 kind::i8: at most 255 windows per sequence), so these are TOP/s"
                bf16, src = measured_tensor_peak()
                row["roofline"] = {"kernel": "syrk_tc_kernel<1, true>", "bound": "tensor", "achieved": row["tensor_tflops"], "peak": 2.0 * bf16,
                                   "unit": "TOP/s", "frac": row["tensor_tflops"] / (2.0 * bf16), "traffic": None,
                                   "peak_source": "2 x the dense bf16 rate (int8 MMAs issue at twice the bf16 rate): " + src,
                                   "note": "bound by the fill of shared memory with operand tiles (ncu: L2 throughput 62 %, profiles/r02_ncu_dense_u8.txt), "
                                           "528 tiles = 1.78 waves at EP300's size"}
            if best is None or row["e2e_s"] < best["e2e_s"]:
                best = row
            del f
        out[tag] = best
    ci = fasta_workload("EP300", 10, 6, device, t=1, approx=True, max_iters=50)
    if ci is not None:
        ci["workload"] = ("EP300, g=10 m=6, approx with the variance test, t=1, max_iters=50: the reference's own acceptance configuration "
                          "(test/run_check.py:45); one virtual stream, its 50 iterations speculated in one launch group")
        ci["combinations_per_s_device"] = ci["combinations_done_this_rank"] / (ci["device_ms"] * 1e-3) if ci["device_ms"] else None
    out["ep300_approx_t1"] = ci
    prot = fasta_workload("1.1", 10, 6, device)
    if prot is not None:
        prot["workload"] = "protein remote homology 1.1 train+test (BASELINE configs[2]), g=10 m=6 exact, 210 combinations, 5 bits per character"
    out["protein_1_1"] = prot
    return out


def fasta_workload(name, g, m, device, dist=None, **kw):
    """One bundled FASTA set through the public API: read, compute_kernel, train kernel to the host."""
    tr, te = os.path.join(ROOT, "data", f"{name}.train.fasta"), os.path.join(ROOT, "data", f"{name}.test.fasta")
    if not (os.path.exists(tr) and os.path.exists(te)):
        return None
    from fastsk_b200 import FastSK, FastaUtility
    fu = FastaUtility()
    Xtr, _ = fu.read_data(tr)
    Xte, _ = fu.read_data(te)
    best = None
    for _ in range(3):
        f = FastSK(g, m, seed=0, profile=True, **({"device": device, "distributed": False} if dist is None else {}), **kw)
        t0 = time.perf_counter()
        f.compute_kernel(Xtr, Xte)
        Ktr = f.get_train_kernel()
        wall = time.perf_counter() - t0
        st = f.stats()
        row = {"e2e_s": wall, "device_ms": st["ms_total"], "combinations_done_this_rank": st["combos_done"],
               "combinations_per_s_e2e_this_rank": st["combos_done"] / wall, "acc_path": st["acc_path"], "n_seq": st["n_seq"],
               "nfeat": st["nfeat"], "key_bits": st["key_bits"], "sort_passes": st["sort_passes"], "kernel_launches": st["kernel_launches"],
               "stdevs": len(f.get_stdevs()), "last_stdev": (f.get_stdevs() or [None])[-1], "trace": float(np.trace(Ktr)),
               "phase_ms": {k[3:]: round(st[k], 3) for k in st if k.startswith("ms_") and k != "ms_total"}}
        if best is None or row["e2e_s"] < best["e2e_s"]:
            best = row
        del f
    return best


def aimed_workload(device, dist=None):
    """BASELINE configs[4]: AImed (55-letter alphabet, 6 bits per character, 60-bit keys, key/value sort), g=20 m=10, approx mode
    with the variance / convergence test, t=20 virtual streams (dealt over the ranks under torchrun), seed 0, at most 50 iterations
    per stream.  COLLECTIVE when dist is given."""
    out = fasta_workload("AImed", 20, 10, device, dist, t=20, approx=True, delta=0.025, max_iters=50)
    if out is not None:
        out["workload"] = "AImed train+test (BASELINE configs[4]), g=20 m=10, approx (variance on), t=20, delta=0.025, max_iters=50, seed=0"
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--combos", type=int, default=0, help="combinations per build (0 = all C(16,8) = 12870; smaller only for quick tests)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison after the timed region (~10 s)")
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
