"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/ncu_launches.py gpurun_out/launches.csv > profiles/launches.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
    tot[r[ik]][0] += 1
    tot[r[ik]][1] += ms
total = sum(v[1] for v in tot.values())
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} launches  total {ms:9.3f} ms  mean {ms / n:8.4f} ms  share {ms / total:6.1%}  {k}")
