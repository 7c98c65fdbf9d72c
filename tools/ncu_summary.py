"""Compact per-kernel summary of an .ncu-rep (one line per distinct kernel/grid)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "t"), ("smsp__inst_executed.sum", "winst"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "l1wf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemwf"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("launch__grid_size", "grid")]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    key = name + r[hdr.index("launch__grid_size")]
    if key in seen:
        continue
    seen.add(key)
    out = [name]
    for c, lab in cols:
        if c in hdr:
            v = r[hdr.index(c)]; u = units[hdr.index(c)]
            try:
                fv = float(v.replace(",", ""))
                v = ("%.3g" % fv)
            except Exception:
                pass
            out.append("%s=%s%s" % (lab, v, u if lab in ("t", "dramR", "dramW") else ""))
    print(" ".join(out))
