"""Regenerate the tracked profile artefacts from scratch captures.

  profile_artifacts.py traffic <full.ncu-rep> <out.json> [frames_in_capture frames_per_step] [--build-id=ID]
      per-stage DRAM bytes and warp instructions of one step (bench.py reads this for roofline.traffic).  When the
      capture covers only part of a step (the device path runs a batch as two halves), pass the frame counts and the
      sums are scaled to a full step.
  profile_artifacts.py launches <launch_list.csv> <out.txt> "<command that was profiled>"
      per-kernel aggregate of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
"""
import collections, csv, io, json, subprocess, sys

STAGE = {"k_level0": "level0", "k_resize": "resize", "k_resize_gather": "resize", "k_fast_cells": "fast",
         "k_octree": "octree", "k_blur": "blur", "k_describe": "describe"}


def stage_of(name):
    base = name.split("(")[0].split("<")[0].split("::")[-1].strip().split(" ")[-1]
    return STAGE.get(base)


def traffic(rep, out, scale=1.0, build_id=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ci = {c: hdr.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum")}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    b, n, w = collections.OrderedDict(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        st = stage_of(r[ci["Kernel Name"]])
        if st is None:
            continue
        tot = 0.0
        for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ci[c]].replace(",", "")) * mult[units[ci[c]]]
        b[st] = b.get(st, 0.0) + tot
        n[st] += 1
        w[st] += float(r[ci["smsp__inst_executed.sum"]].replace(",", ""))
    doc = {"source": "%s (ncu --set full --clock-control none); dram__bytes_read.sum + dram__bytes_write.sum and "
                     "smsp__inst_executed.sum summed over the stage's launches, scaled x%.3g to one full step" % (rep, scale),
           "build_id": build_id,   # orbx_build_id() of the library that was profiled (printed by tools/prof_step.py)
           "bytes_per_step": {k: v * scale for k, v in b.items()},
           "launches": {k: int(v) for k, v in n.items()},
           "warp_instructions_per_step": {k: v * scale for k, v in w.items()}}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


def launches(csv_path, out, cmd):
    text = open(csv_path, errors="ignore").read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        name = r["Kernel Name"].split("(")[0][:44]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none: %s" % cmd,
             "per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  Full CSV: %s (scratch)" % csv_path, ""]
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-44s n=%4d total=%10.1f us avg=%9.1f us share=%5.1f%%" % (name, cnt, us, us / cnt, 100 * us / tot))
    ext = {k: v for k, v in agg.items() if stage_of(k)}
    et = sum(v[1] for v in ext.values()) or 1.0
    lines += ["", "extraction stages only (share of the extraction kernels' time):"]
    for name, (cnt, us) in sorted(ext.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-44s share=%5.1f%%" % (name, 100 * us / et))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        args = [a for a in sys.argv[2:] if not a.startswith("--build-id=")]
        bid = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--build-id=")), None)
        sc = float(args[3]) / float(args[2]) if len(args) > 3 else 1.0
        traffic(args[0], args[1], sc, bid)
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
