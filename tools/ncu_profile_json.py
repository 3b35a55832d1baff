"""The numbers bench.py quotes from an ncu --set full capture of the dominant kernel (one launch), as JSON.
    python tools/ncu_profile_json.py gpurun_out/prof.ncu-rep c3 profiles/r01_c3_step13_current.txt > profiles/r01_c3_profile.json"""
import csv
import io
import json
import subprocess
import sys

rep, workload, source = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(u, 1.0)
    return v * scale


out = {
    "workload": workload, "kernel": vals[hdr.index("Kernel Name")],
    "source": f"{source} (ncu --set full --clock-control none, one launch)",
    "kernel_ms_under_ncu": round(get("gpu__time_duration.sum"), 4),
    "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
    "dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
    "issue_active_pct": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
    "pipe_fma_pct": round(get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
    "pipe_alu_pct": round(get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
    "warp_instructions": get("smsp__inst_executed.sum"),
    "threads_per_instruction": round(get("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
    "registers_per_thread": get("launch__registers_per_thread"),
    "warps_active_pct": round(get("sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
}
# FP32 work the hardware executed: predicated-on thread instructions per opcode from the SASS page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(srows) if "Source" in r and "Instructions Executed" in r)
h = srows[hi]
isrc, ith = h.index("Source"), h.index("Predicated-On Thread Instructions Executed")
ops = {}
for r in srows[hi + 1:]:
    try:
        parts = r[isrc].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ith])
    except (IndexError, ValueError):
        pass
out["thread_fadd"], out["thread_fmul"], out["thread_ffma"] = ops.get("FADD", 0), ops.get("FMUL", 0), ops.get("FFMA", 0)
out["executed_fp32_flops"] = ops.get("FADD", 0) + ops.get("FMUL", 0) + 2 * ops.get("FFMA", 0) + ops.get("FMNMX", 0) + ops.get("MUFU", 0)
print(json.dumps(out, indent=1))
