"""Print the hottest SASS regions (by executed warp instructions) of an .ncu-rep."""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr_frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0012
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ia]), r[isrc]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
hot = [i for i, d in enumerate(data) if d[0] > tot * thr_frac]
print(len(hot), "hot SASS lines cover", round(sum(data[i][0] for i in hot) / tot, 3), "of", tot, "warp instructions")
prev = None
for i in hot:
    if prev is not None and i != prev + 1:
        print("   ...")
    print(f"{i:6d} {data[i][0] / 1e6:8.1f}M  {data[i][1]}")
    prev = i
