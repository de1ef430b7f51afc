#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per launch the duration, DRAM bytes, pipe utilisation and top stall reasons.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
BASIC = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
         "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
         "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
         "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
for d in data:
    print("=" * 100)
    print(d[idx["Kernel Name"]][:160])
    for k in BASIC:
        if k in idx:
            print(f"  {k:88s} {d[idx[k]]:>18s} {units[idx[k]]}")
    pipes = []
    stalls = []
    for h in hdr:
        try:
            v = float(d[idx[h]])
        except ValueError:
            continue
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active"):
            pipes.append((v, h))
        if "average_warp" in h and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") or \
           (h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio")):
            stalls.append((v, h))
    for v, h in sorted(pipes, reverse=True)[:5]:
        print(f"  pipe  {h:82s} {v:18.2f} %")
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {h:82s} {v:18.2f}")
