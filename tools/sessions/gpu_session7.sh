#!/bin/bash
# GPU session 7: first run of the dense tensor-core path (tcgen05 + TMA): diagnostic, its parity tests, then the whole suite
mkdir -p gpurun_out
timeout 300 python tools/dense_check.py > gpurun_out/s7_dense_check.txt 2>&1; echo "dense_check rc=$?"
tail -25 gpurun_out/s7_dense_check.txt
( time timeout 600 python -m pytest tests -m gpu -x -q -k "dense" ) > gpurun_out/s7_pytest_dense.txt 2>&1
tail -15 gpurun_out/s7_pytest_dense.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s7_pytest.txt 2>&1
tail -8 gpurun_out/s7_pytest.txt
timeout 600 python tools/run_configs.py > gpurun_out/s7_configs.jsonl 2> gpurun_out/s7_configs.err
cut -c1-330 gpurun_out/s7_configs.jsonl
