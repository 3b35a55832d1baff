"""Attribute the executed instructions of a profiled kernel to SOURCE LINES.
    python tools/ncu_lines.py <report.ncu-rep> <object-with-cubin (.o / .so)> <kernel symbol substring> [top N]
ncu's SASS page gives executed warp instructions per instruction address; nvdisasm -g gives address -> file:line.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubins = sorted((os.path.getsize(os.path.join(tmp, f)), f) for f in os.listdir(tmp))
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubins[-1][1])], capture_output=True, text=True).stdout
line_of, cur, active = {}, ("?", 0), False
for ln in dis.splitlines():
    if ln.lstrip().startswith(".section"):
        active = ".text." in ln and sym in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, iaddr, ist = hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("# Samples")
ith = hdr.index("Thread Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[iaddr], 16) if r[iaddr].startswith("0x") else int(r[iaddr]), int(r[ia]), int(r[ist] or 0), int(r[ith] or 0)))
    except (ValueError, IndexError):
        pass
base = min(d[0] for d in data)
by_line = collections.Counter()
samples = collections.Counter()
n_inst = collections.Counter()
threads = collections.Counter()
for addr, n, st, th in data:
    key = line_of.get(addr - base, (("?", 0), ""))[0]
    by_line[key] += n
    samples[key] += st
    n_inst[key] += 1
    threads[key] += th
tot, tots = sum(by_line.values()), max(sum(samples.values()), 1)
print(f"{tot} warp instructions, {len(data)} SASS instructions, {len(by_line)} source lines")
srcs = {}
for (f, l), n in by_line.most_common(top):
    if f not in srcs:
        for root in ("ray_tracer_challenge_b200/csrc", "."):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().splitlines()
                break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{n / tot:6.2%} instr {samples[(f, l)] / tots:6.2%} stall  {threads[(f, l)] / max(n, 1):5.1f} lanes {n_inst[(f, l)]:4d} sass  {f}:{l:<5d} {text}")

# ---- the same, grouped by the device function a line belongs to
print("\nby function:")
fn_start = {}
for f, lines in srcs.items():
    starts = []
    for i, text in enumerate(lines):
        m = re.match(r"\s*(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", text)
        if m and not text.strip().startswith("//"):
            starts.append((i + 1, m.group(1)))
    fn_start[f] = starts
by_fn, st_fn, th_fn = collections.Counter(), collections.Counter(), collections.Counter()
for (f, l), n in by_line.items():
    if f not in srcs:
        for root in ("ray_tracer_challenge_b200/csrc", "."):
            pth = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(pth):
                srcs[f] = open(pth).read().splitlines()
                break
        else:
            srcs[f] = []
        starts = []
        for i, text in enumerate(srcs[f]):
            m = re.match(r"\s*(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", text)
            if m and not text.strip().startswith("//"):
                starts.append((i + 1, m.group(1)))
        fn_start[f] = starts
    name = f
    for s0, nm in fn_start.get(f, []):
        if s0 <= l:
            name = nm
    by_fn[name] += n
    st_fn[name] += samples[(f, l)]
    th_fn[name] += threads[(f, l)]
for name, n in by_fn.most_common(30):
    print(f"{n / tot:6.2%} instr {st_fn[name] / tots:6.2%} stall  {th_fn[name] / max(n, 1):5.1f} lanes  {name}")
