"""Hot source lines of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_source.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci, csmp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.OrderedDict()
cur = None
for r in rows[hi + 1:]:
    if len(r) < 4:
        continue
    if r[0] != "":
        cur = (r[0], r[1].strip()); agg.setdefault(cur, [0, 0]); continue
    if cur is None or r[2] in ("", "..."):
        continue
    try:
        agg[cur][0] += int(r[ci]); agg[cur][1] += int(r[csmp] or 0)
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print("kernel %s: %d warp-instructions, %d samples" % (kern, tot, tots))
for (ln, src), (n, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  L%-4s %s" % (100.0 * n / tot, 100.0 * smp / tots, ln, src[:120]))
