#!/bin/bash
for v in 0 1 2; do echo "== ORBX_OCT_BIG=$v"; ORBX_OCT_BIG=$v python tools/prof_step.py --workload 4k --warm 2 --steps 5 --split 1 2>&1 | grep -E "stage ms|ms per step"; done
python tools/prof_step.py --warm 3 --steps 20 --split 1 2>&1 | grep -E "stage ms|ms per step"
python tools/prof_step.py --workload kitti --warm 3 --steps 10 --split 1 2>&1 | grep -E "stage ms|ms per step"
python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_golden.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | tail -2
ORBX_OCT_BIG=1 python -m pytest tests/test_gpu_extract_parity.py -m gpu -q -x -k 4k 2>&1 | tail -2
