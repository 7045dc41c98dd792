"""Summarise an .ncu-rep here (no GPU): key raw metrics + per-source-line instruction/stall shares.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("  %-85s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
agg = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iA, iI, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or r[iA] != "-":
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[iI]), int(r[iS])
    except ValueError:
        continue
    a = agg.setdefault(ln, [0, 0, r[1]])
    a[0] += inst
    a[1] += samp
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (ti, ts))
for ln, (inst, samp, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp  L%-4d %s" % (100 * inst / ti, 100 * samp / ts, ln, text.strip()[:105]))
