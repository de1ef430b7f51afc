#!/bin/bash
# GPU session 4: parity, accumulate unroll / pipelining knobs, segment cost decomposition, full ncu capture of segment + accumulate
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.txt 2>&1
tail -3 gpurun_out/s4_pytest.txt
rm -f gpurun_out/s4_steps.txt
for opts in "" "--acc-unroll 2" "--acc-unroll 6" "--acc-unroll 8" "--acc-pipe 1 --acc-unroll 2" "--acc-pipe 1 --acc-unroll 3" "--acc-pipe 1 --acc-unroll 4" \
            "--seg-exp 1" "--seg-exp 2" "--seg-exp 3"; do
  echo "== $opts" >> gpurun_out/s4_steps.txt
  timeout 300 python tools/profile_step.py --batch 48 --reps 2 $opts 2>&1 | head -1 >> gpurun_out/s4_steps.txt
done
cat gpurun_out/s4_steps.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"segment_kernel|accumulate_rows" -c 2 -o gpurun_out/s4_seg_acc \
    python tools/profile_step.py --batch 8 --wave 400 --reps 0 > gpurun_out/s4_ncu.log 2>&1
tail -2 gpurun_out/s4_ncu.log
