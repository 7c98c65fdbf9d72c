#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matcher_dropin.py -m gpu -q > gpurun_out/r02_gputest2.log 2>&1; tail -3 gpurun_out/r02_gputest2.log
python tools/prof_step.py --warm 2 --steps 10 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 24 -c 12 -o gpurun_out/r02_full_base -f python tools/prof_step.py --warm 2 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
