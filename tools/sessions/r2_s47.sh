#!/bin/bash
# r2 session 47 (1 GPU): the round-end sequence on the final code: smoke(), the full GPU suite
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s47_smoke.txt 2>&1; tail -1 gpurun_out/r2s47_smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s47_pytest.txt 2>&1
tail -2 gpurun_out/r2s47_pytest.txt
