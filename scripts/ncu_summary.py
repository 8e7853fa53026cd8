"""Text summaries of an .ncu-rep for profiles/:  python scripts/ncu_summary.py REP OUT_PREFIX [title]
writes OUT_PREFIX_ncu_raw.csv (every raw metric, one column per captured launch) and OUT_PREFIX_stalls.txt
(warp stall samples + the headline counters of each launch)."""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
with open(out + "_ncu_raw.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[ki][:64] for r in launches])
    for i, h in enumerate(hdr):
        w.writerow([h, units[i]] + [r[i] for r in launches])
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
with open(out + "_stalls.txt", "w") as f:
    f.write(title + "\n")
    for r in launches:
        f.write("\n== %s\n" % r[ki][:110])
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write("%-80s %-14s %s\n" % (k, units[i], r[i]))
        st = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1
        f.write("warp stall samples:\n")
        for h, v in sorted(st, key=lambda x: -x[1]):
            if v: f.write("  %-34s %8d %5.1f%%\n" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v, 100 * v / tot))
print(open(out + "_stalls.txt").read())
