"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`): key utilisation metrics, stall reasons, SASS mix.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/summary.txt]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg",
]
for ridx, vals in enumerate(rows[2:]):
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else f"kernel {ridx}"
    print(f"== {name}")
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            print(f"  {h:76s} {v:>18s} {u}")
    print("  -- stall reasons (warps per issue-active cycle)")
    st = [(float(v), h) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
byop, samp, tot, nl = collections.Counter(), collections.Counter(), 0, 0
for r in rows[hi + 1:]:
    if len(r) <= ia:
        continue
    try:
        n, s = int(r[ia]), int(r[isamp])
    except ValueError:
        continue
    parts = r[isrc].split()
    if not parts:
        continue
    op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
    byop[op] += n
    samp[op] += s
    tot += n
    nl += 1
print(f"  -- SASS mix ({nl} SASS lines, {tot} warp instructions)")
ts = max(1, sum(samp.values()))
for op, n in byop.most_common(24):
    print(f"  {op:10s} {n / tot * 100:6.2f}% of instructions  {samp[op] / ts * 100:6.2f}% of stall samples")
