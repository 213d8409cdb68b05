"""Per-source-line stall summary of an ncu capture (--set full --import-source on, -lineinfo build):
python tools/ncu_line_summary.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
print(v[h.index("Kernel Name")] if "Kernel Name" in h else "")
for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max"):
    if k in h:
        print("  %-70s %s %s" % (k, v[h.index(k)], u[h.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if "# Samples" in x)
hdr = rows[hi]
ixs = hdr.index("# Samples")
stalls = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
fname = ""
lines = []
for x in rows:
    if len(x) == 2 and x[0] == "File Path":
        fname = x[1].split("/")[-1]
    if len(x) == len(hdr) and x[0].strip().isdigit() and x[ixs].isdigit():
        lines.append((fname, x))
tot = sum(int(x[ixs]) for _, x in lines)
agg = {}
for _, x in lines:
    for i, s in stalls:
        if x[i].isdigit():
            agg[s] = agg.get(s, 0) + int(x[i])
print("samples", tot, "stall mix:", ", ".join("%s %.1f%%" % (k, 100 * n / tot) for k, n in sorted(agg.items(), key=lambda t: -t[1])[:9]))
for f, x in lines:
    n = int(x[ixs])
    if n > tot * thr / 100:
        st = sorted(((int(x[i]) if x[i].isdigit() else 0, s) for i, s in stalls), reverse=True)[:3]
        print("%-22s %5s %5.1f%%  %-95s %s" % (f[:22], x[0], 100 * n / tot, x[1].strip()[:95], " ".join("%s:%.0f%%" % (b, 100 * a / max(n, 1)) for a, b in st)))
