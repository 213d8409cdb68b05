"""Extract the judged metrics of every kernel in an .ncu-rep into a small text/JSON summary."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]
idx = [(w, h.index(w)) for w in want if w in h]
res = []
for r in rows[2:]:
    d = {w: r[i] for w, i in idx}
    d["units"] = {w: rows[1][i] for w, i in idx if rows[1][i]}
    res.append(d)
print(json.dumps(res, indent=1))
