#!/bin/bash
# round-2 evidence on the shipped objects: bench line, ncu launch list of the same command, ncu --set full of one 256-frame step
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-latency --skip-extra > /dev/null 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-latency --skip-extra > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/prof_step.py --warm 2 --steps 10 > gpurun_out/r02_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -s 24 -c 12 -o gpurun_out/r02_full_final -f python tools/prof_step.py --warm 2 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
cat gpurun_out/r02_prof_plain.log
python -c "
import numpy as np, sys
sys.path.insert(0,'.')
from oracle import oracle as O
from orbslam_in_practice_b200.synth import synth_batch
ex = O.OracleExtractor(1000,1.2,8,20,7)
r=np.zeros(8,int); c=np.zeros(8,int); k=0
for f in range(32):
    ko,_ = ex(synth_batch([f])[0]); k+=len(ko)
    r += [ex.retries(l) for l in range(8)]; c += [len(ex.candidates(l)) for l in range(8)]
print('synth.py seeds 0..31: mean keypoints %.1f, min-threshold retry cells per level (mean)'%(k/32), (r/32).round(1).tolist(), 'candidates per level (mean)', (c/32).round(0).tolist())
" > gpurun_out/r02_synth_coverage.txt 2>&1; cat gpurun_out/r02_synth_coverage.txt
python tools/latency_probe.py > gpurun_out/r02_latency_final.log 2>&1; tail -4 gpurun_out/r02_latency_final.log
