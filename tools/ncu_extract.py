"""Condense `ncu --page raw --csv` exports into the handful of counters DESIGN.md quotes (one row per captured launch).
python tools/ncu_extract.py gpurun_out/x_raw.csv > profiles/ncu_r02_x.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
name_i = hdr.index("Kernel Name")
cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
extra = [h for h in hdr if "fp64" in h and h not in WANT][:6]
cols += [(h, hdr.index(h)) for h in extra]
out = csv.writer(sys.stdout)
out.writerow(["kernel"] + [f"{w} [{units[i]}]" for w, i in cols])
for r in rows[2:]:
    if len(r) <= name_i:
        continue
    out.writerow([r[name_i][:60]] + [r[i] for _, i in cols])
