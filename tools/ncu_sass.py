"""SASS listing of one kernel from an .ncu-rep with per-instruction executed counts and stall samples.
usage: ncu_sass.py report.ncu-rep kernel_regex > out.txt"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
cw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
tot = sum(int(r[ci]) for r in rows[hi + 1:] if len(r) > ci and r[ci].isdigit()) or 1
acc = 0
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= ci or not r[ci].isdigit():
        continue
    n = int(r[ci]); acc += n
    print("%4d %5.2f%% %6.2f%% %9d thr=%-4s smp=%-5s wf=%-8s %s" % (i, 100.0 * n / tot, 100.0 * acc / tot, n, r[ct], r[cs], r[cw] if cw else "", r[1].strip()))
print("total warp instructions:", tot)
