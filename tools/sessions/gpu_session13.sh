#!/bin/bash
# GPU session 13: ncu of the Welford step (variance mode, EP300 shape)
mkdir -p gpurun_out
timeout 300 python tools/approx_step.py
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"welford_kernel" -s 4 -c 1 -o gpurun_out/s13_welford \
    python tools/approx_step.py --iters 2 > gpurun_out/s13_ncu.log 2>&1
tail -2 gpurun_out/s13_ncu.log
