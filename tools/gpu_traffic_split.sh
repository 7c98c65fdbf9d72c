#!/bin/bash
# DRAM traffic and time of one 256-frame VGA step against the number of device sub-batches (the working set of a
# sub-batch -- pyramid + blurred levels, 2.1 MB per frame -- fits the 126 MB L2 from 8 sub-batches on).
# ncu with --cache-control none so that the L2 keeps what the previous kernel left, one pass (two DRAM counters).
mkdir -p gpurun_out
for sp in 1 2 4 8 16; do
  python tools/prof_step.py --warm 3 --steps 20 --split $sp 2>&1 | grep "ms per step"
  n=$((12 * sp))
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum -s $((2 * n)) -c $n --csv --log-file gpurun_out/traffic_split_$sp.csv python tools/prof_step.py --warm 2 --split $sp > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/traffic_split_$sp.csv")) if len(r) > 5 and r[0].strip('"').isdigit()]
tot = 0.0
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows:
    tot += float(r[-1].replace(",", "")) * mult.get(r[-2], 1.0)
print("split $sp: %d launches, DRAM traffic %.1f MB per step = %.2f MB per frame" % (len(rows) // 2, tot / 1e6, tot / 1e6 / 256))
PY
done
