#!/bin/bash
# quick GPU check of a kernel change: extraction parity tests, then stage times of the 256-frame VGA step
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python tools/prof_step.py --warm 3 --steps 20 --split 1 > gpurun_out/${TAG}_prof.log 2>&1; cat gpurun_out/${TAG}_prof.log
python tools/prof_step.py --warm 3 --steps 20 --split 2 2>&1 | grep "ms per step"
