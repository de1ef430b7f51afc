#!/usr/bin/env python
"""Stall samples per CUDA source line of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
    python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rd = csv.reader(io.StringIO("\n".join(lines[start:])))
hdr = next(rd)
si = hdr.index("# Samples")
rows = [r for r in rd if len(r) > si and r[0].strip().isdigit() and r[si].isdigit()]
tot = sum(int(r[si]) for r in rows)
print(f"samples {tot}")
for r in sorted(rows, key=lambda r: -int(r[si]))[:top]:
    print(f"{100 * int(r[si]) / tot:5.1f}%  L{r[0]:>4s}  {r[1].strip()[:140]}")
