#!/bin/bash
# ncu --set full of one 4K step (8 frames): octree + FAST source pages
mkdir -p gpurun_out
python tools/prof_step.py --workload 4k --frames 8 --warm 2 --steps 5 > gpurun_out/prof4k_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_octree|k_fast" -s 4 -c 2 -o gpurun_out/r02_full_4k -f python tools/prof_step.py --workload 4k --frames 8 --warm 2 > gpurun_out/ncu4k.log 2>&1
cat gpurun_out/prof4k_plain.log; tail -3 gpurun_out/ncu4k.log
