"""Linear listing of a profiled kernel's executed SASS with stall samples, grouped into regions of consecutive
instructions: where in program order the warps spend their time, and on which stall reason.
    python tools/ncu_regions.py <report.ncu-rep | source.csv> [min share of samples per region, default 0.004]
"""
import csv
import io
import subprocess
import sys

src = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
text = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hi + 1:]:
    try:
        ex = int(r[col["Instructions Executed"]])
        smp = int(r[col["# Samples"]] or 0)
        st = {n: int(r[col[n]] or 0) for n in reasons}
        data.append((r[col["Source"]], ex, smp, st))
    except (ValueError, IndexError):
        pass
tot_ex = sum(d[1] for d in data)
tot_s = sum(d[2] for d in data)
print(f"{len(data)} SASS, {tot_ex} warp instr, {tot_s} samples")
# regions: maximal runs of executed instructions; split at BRA/EXIT/BSYNC/CALL/RET boundaries with count change > 2x
regions, cur = [], []
for i, d in enumerate(data):
    if d[1] == 0:
        if cur:
            regions.append(cur)
            cur = []
        continue
    if cur and (d[1] > 2 * data[cur[-1]][1] or 2 * d[1] < data[cur[-1]][1]):
        regions.append(cur)
        cur = []
    cur.append(i)
if cur:
    regions.append(cur)
for reg in regions:
    s = sum(data[i][2] for i in reg)
    if s < thr * tot_s:
        continue
    ex = sum(data[i][1] for i in reg)
    agg = {}
    for i in reg:
        for k, v in data[i][3].items():
            agg[k] = agg.get(k, 0) + v
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    print(f"\n== SASS {reg[0]}..{reg[-1]} ({len(reg)} instr)  x{data[reg[0]][1] / 1e3:.0f}k  instr {100 * ex / tot_ex:.2f}%  samples {100 * s / tot_s:.2f}%  "
          + "  ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in top))
    # the instructions that collect the samples
    worst = sorted(reg, key=lambda i: -data[i][2])[:6]
    for i in sorted(worst):
        if data[i][2] * 50 < s:
            continue
        r = max(data[i][3].items(), key=lambda kv: kv[1])
        print(f"     {i:6d} {100 * data[i][2] / tot_s:5.2f}%  {r[0][6:]:<18} {data[i][0][:100]}")
