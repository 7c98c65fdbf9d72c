#!/bin/bash
mkdir -p gpurun_out
for st in 0 1 2; do for cv in 72 100; do
  echo "== stage $st carveout $cv"
  ORBX_DESC_STAGE=$st ORBX_DESC_CARVEOUT=$cv python tools/prof_step.py --warm 3 --steps 20 --split 1 2>&1 | grep -E "stage ms|ms per step"
done; done
for st in 1 2; do ORBX_DESC_STAGE=$st python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -2; done
