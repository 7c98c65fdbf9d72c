#!/bin/bash
# quick GPU check of a kernel change: extraction parity tests, then stage times of the 256-frame VGA step (+ KITTI / 4K with ALL=1)
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_golden.py tests/test_gpu_baseline_sizes.py -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python tools/prof_step.py --warm 3 --steps 20 --split 1 > gpurun_out/${TAG}_prof.log 2>&1; cat gpurun_out/${TAG}_prof.log
python tools/prof_step.py --warm 3 --steps 20 --split 2 2>&1 | grep "ms per step"
if [ -n "$ALL" ]; then
  python tools/prof_step.py --workload kitti --warm 3 --steps 10 --split 1 2>&1 | tee gpurun_out/${TAG}_prof_kitti.log
  python tools/prof_step.py --workload 4k --warm 2 --steps 5 --split 1 2>&1 | tee gpurun_out/${TAG}_prof_4k.log
fi
if [ -n "$LAT" ]; then python tools/latency_probe.py 2>&1 | tee gpurun_out/${TAG}_latency.log; fi
