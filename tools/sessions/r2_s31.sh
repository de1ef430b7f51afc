#!/bin/bash
# r2 session 31 (8 GPUs): device -> host bandwidth of all GPUs into one host buffer
mkdir -p gpurun_out
numactl -H > gpurun_out/r2s31_numa.txt 2>&1 || ls /sys/devices/system/node/ > gpurun_out/r2s31_numa.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2s31_numa.txt 2>&1
FSK_TRACE=1 timeout 300 python tools/d2h_bw.py > gpurun_out/r2s31_d2h.txt 2>&1
cat gpurun_out/r2s31_d2h.txt; head -30 gpurun_out/r2s31_numa.txt
