"""Per-phase summary of a band-reduction ncu capture (--set full --import-source on): splits the
SASS of the kernel at its block barriers, and for every DMMA-carrying segment reports its share of
the warp-stall samples, the tensor-pipe efficiency (DMMA issued x 16 cycles per scheduler against
the segment's elapsed cycles) and the leading stall reasons; then the hottest non-DMMA
instructions.        python tools/ncu_phase_summary.py report.ncu-rep [kernel-regex] > summary.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[2]
def g(k):
    return v[h.index(k)] if k in h else "n/a"
cyc = float(g("sm__cycles_elapsed.max").replace(",", ""))
grid = int(float(g("launch__grid_size").replace(",", "")))
print("duration_ms", g("gpu__time_duration.sum"), "cycles", cyc, "grid", grid, "regs", g("launch__registers_per_thread"))
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"):
    print(k, g(k), r[1][h.index(k)] if k in h else "")
for k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"):
    print(k, g(k))
stalls = [s for s in hdr if s.startswith("stall_") and "Not Issued" not in s]
tot = sum(int(x[ix["# Samples"]]) for x in data)
agg = {s[6:]: sum(int(x[ix[s]]) for x in data) for s in stalls}
print("samples", tot, "stall mix:", ", ".join("%s %.1f%%" % (k, 100 * n / tot) for k, n in sorted(agg.items(), key=lambda t: -t[1])[:8]))
sass = {}
for x in data:
    op = x[1].split()[0] if x[1].split() else ""
    if op.startswith("@"):
        op = x[1].split()[1]
    for key in ("UBLKCP", "SYNCS", "LDGSTS", "DMMA", "UTMALDG", "BAR", "STL", "LDL"):
        if op.startswith(key):
            sass[key] = sass.get(key, 0) + 1
print("SASS instruction counts in this kernel:", sass)
seg, cur, curd, start = [], 0, 0, 0
for i, x in enumerate(data):
    cur += int(x[ix["# Samples"]])
    curd += "DMMA" in x[1]
    if "BAR.SYNC" in x[1] or i == len(data) - 1:
        seg.append((start, i, cur, curd)); cur = curd = 0; start = i + 1
print("\nsegments between block barriers (>= 0.5 % of the samples):")
for a, b, n, nd in seg:
    if n < 0.005 * tot:
        continue
    dm = sum(int(x[ix["Instructions Executed"]]) for x in data[a:b + 1] if "DMMA" in x[1])
    t = n / tot * cyc
    ideal = dm / grid * 16 / 4
    st = {s[6:]: sum(int(x[ix[s]]) for x in data[a:b + 1]) for s in stalls}
    top = ", ".join("%s %.0f%%" % (k, 100 * c / n) for k, c in sorted(st.items(), key=lambda q: -q[1])[:4])
    print("  sass %5d-%5d  %5.1f %% of samples  dmma_sass %3d  tensor-pipe efficiency %4.2f   %s"
          % (a, b, 100 * n / tot, nd, ideal / t if t else 0.0, top))
print("\nhottest non-DMMA instructions:")
oth = sorted(((int(x[ix["# Samples"]]), i, x[1].strip()[:64]) for i, x in enumerate(data)
              if "DMMA" not in x[1] and x[1].strip() != "NOP"), reverse=True)[:14]
for n, i, t in oth:
    x = data[i]
    top = sorted(((int(x[ix[s]]), s[6:]) for s in stalls), reverse=True)[0]
    print("  %6d (%4.1f %%)  sass %5d  %-64s %s" % (n, 100 * n / tot, i, t, top[1]))
