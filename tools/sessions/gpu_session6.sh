#!/bin/bash
# GPU session 6: state check after the container was re-created: parity suite, default bench, bundled configurations
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_pytest.txt 2>&1
tail -3 gpurun_out/s6_pytest.txt
timeout 600 python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
tail -c 600 gpurun_out/s6_bench.json
timeout 600 python tools/run_configs.py > gpurun_out/s6_configs.jsonl 2> gpurun_out/s6_configs.err
cat gpurun_out/s6_configs.jsonl | cut -c1-420
