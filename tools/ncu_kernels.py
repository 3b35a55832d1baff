"""Per-kernel totals of an ncu CSV launch list (--metrics gpu__time_duration.sum,smsp__inst_executed.sum,
smsp__thread_inst_executed_per_inst_executed.ratio): launches, time, warp instructions, lanes per instruction.
    python tools/ncu_kernels.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
un = h.index("Metric Unit")
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    if r[mn] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[un], 1e-6)
    per.setdefault(r[idc], {"k": r[kn]})[r[mn]] = v
agg = collections.OrderedDict()
for d in per.values():
    name = d["k"].split("(")[0].replace("void ", "").replace("rtc::strict::", "").replace("rtc::fast::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    inst = d.get("smsp__inst_executed.sum", 0.0)
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += inst
    a[3] += inst * d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
total = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{k:42s} launches {a[0]:4d}  {a[1]:9.3f} ms ({100 * a[1] / max(total, 1e-9):5.1f} %)  warp-instr {a[2] / 1e6:9.1f} M  lanes/instr {a[3] / max(a[2], 1):5.1f}")
print(f"{'total':42s} {total:9.3f} ms")
