#!/bin/bash
# round-2 GPU session 1: full GPU test suite, TMA probes, baseline bench + latency
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_box.txt 2>&1
nvidia-smi -q | grep -i -E "MIG|Confidential|Virtualization|Persistence" >> gpurun_out/r02_box.txt 2>&1
nproc >> gpurun_out/r02_box.txt; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/r02_box.txt
for m in bulk tensor prefetch; do timeout 60 tools/_build/tma_probe3 $m >> gpurun_out/r02_tma_probe.log 2>&1; echo "exit=$?" >> gpurun_out/r02_tma_probe.log; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputest1.log
timeout 600 python bench.py > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err; echo "rc=$?" >> gpurun_out/r02_bench_base.err
timeout 300 python tools/latency_probe.py > gpurun_out/r02_latency_base.log 2>&1
tail -5 gpurun_out/r02_gputest1.log; cat gpurun_out/r02_tma_probe.log; cat gpurun_out/r02_latency_base.log
