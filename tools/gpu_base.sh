#!/bin/bash
# round-2 baseline GPU session: box facts, TMA probes, full GPU suite, bench, latency, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_box.txt 2>&1
nvidia-smi -q | grep -i -E "MIG|Confidential|Virtualization|Persistence" >> gpurun_out/r02_box.txt 2>&1
nproc >> gpurun_out/r02_box.txt; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/r02_box.txt
nvidia-smi topo -m >> gpurun_out/r02_box.txt 2>&1
for m in bulk tensor prefetch; do echo "== $m" >> gpurun_out/r02_tma_probe.log; timeout 60 tools/_build/tma_probe3 $m >> gpurun_out/r02_tma_probe.log 2>&1; echo "exit=$?" >> gpurun_out/r02_tma_probe.log; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputest1.log
timeout 600 python bench.py > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err; echo "rc=$?" >> gpurun_out/r02_bench_base.err
timeout 300 python tools/latency_probe.py > gpurun_out/r02_latency_base.log 2>&1
python tools/prof_step.py --warm 2 --steps 10 > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -s 24 -c 12 -o gpurun_out/r02_full_base -f python tools/prof_step.py --warm 2 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/r02_gputest1.log; cat gpurun_out/r02_tma_probe.log; cat gpurun_out/r02_latency_base.log; cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
