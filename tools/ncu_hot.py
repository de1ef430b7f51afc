#!/usr/bin/env python
"""Hottest SASS instructions (stall samples) and shared-memory wavefronts per instruction of one kernel in an .ncu-rep.
    python tools/ncu_hot.py report.ncu-rep kernel_regex [launch_index]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r["# Samples"] or "0").isdigit()]
tot = sum(int(r["# Samples"] or 0) for r in rows)
wf = sum(int(r["L1 Wavefronts Shared"] or 0) for r in rows)
print(f"samples {tot}, shared wavefronts {wf}, instructions {len(rows)}")
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:28]:
    st = {k[6:]: int(v) for k, v in r.items() if k and k.startswith("stall_") and "Not Issued" not in k and v and v.isdigit() and int(v) > 0}
    top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*int(r['# Samples'])/tot:5.1f}%  wf={int(r['L1 Wavefronts Shared'] or 0):>10d}  {r['Source'][:70]:70s} {top}")
