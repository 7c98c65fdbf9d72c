#!/bin/bash
# ncu --set full of selected kernels of one 256-frame VGA step: tools/gpu_ncu_kernels.sh TAG 'regex' [count]
mkdir -p gpurun_out
TAG=${1:-k}; RE=${2:-k_fast_cells}; CNT=${3:-1}
python tools/prof_step.py ${WORKLOAD:+--workload $WORKLOAD} --warm 2 --steps 5 > gpurun_out/${TAG}_plain.log 2>&1 && cat gpurun_out/${TAG}_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $((3*CNT)) -c $CNT -o gpurun_out/${TAG} -f python tools/prof_step.py ${WORKLOAD:+--workload $WORKLOAD} --warm 3 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
