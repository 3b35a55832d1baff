"""Instruction-footprint view of a profiled kernel: which SASS instructions one launch EXECUTED, how many KB they span,
which call sites they belong to, and how often the IEEE division / sqrt slow paths ran.
    python tools/ncu_footprint.py <report.ncu-rep> <object-with-cubin (.o)> <kernel symbol substring> [call-site depth]
ncu's source page gives executed warp instructions per instruction address; `nvdisasm -gi` gives, per address, the source
line and the chain of call sites it was inlined through.  Out-of-line device functions live in the kernel's own .text
section, so one section holds everything the kernel can execute.
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
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2

# ---- executed warp / thread instructions per address (relative to the kernel's first instruction)
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(page)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, iaddr, ith = hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Thread Instructions Executed")
raw = []
for r in rows[hi + 1:]:
    try:
        raw.append((int(r[iaddr], 16) if r[iaddr].lower().startswith("0x") else int(r[iaddr]), int(r[ia]), int(r[ith])))
    except (ValueError, IndexError):
        pass
base = min(a for a, _, _ in raw)
executed = {a - base: (n, t) for a, n, t in raw}
total = sum(n for n, _ in executed.values())

# ---- address -> (call-site chain, outermost first), instruction text, enclosing internal subroutine
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
chain_of, text_of, sub_of = {}, {}, {}
active, pending, chain, label = False, [], (), None
for ln in dis.splitlines():
    if ln.lstrip().startswith(".section"):
        active, label = ".text." in ln and sym in ln, None
        continue
    if not active:
        continue
    m = re.match(r"^(\$__internal_\d+_\$\S+):", ln)
    if m:
        label = re.sub(r"^\$__internal_\d+_\$", "", m.group(1))
        continue
    m = re.search(r'//## File "[^"]*/([^"/]+)", line (\d+)', ln)
    if m:
        pending.append(f"{m.group(1)}:{m.group(2)}")
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if pending:
            chain, pending = tuple(reversed(pending)), []
        a = int(m.group(1), 16)
        chain_of[a], text_of[a], sub_of[a] = chain, m.group(2).strip(), label

hot = sorted((n for n, _ in executed.values() if n > 0), reverse=True)
lines128 = {a // 128 for a, (n, _) in executed.items() if n > 0}
print(f"{len(executed)} SASS instructions ({len(executed) * 16 / 1024:.0f} KB), {len(hot)} executed ({len(hot) * 16 / 1024:.1f} KB; "
      f"{len(lines128) * 128 / 1024:.1f} KB in 128-byte lines), {total} warp instructions")
for frac in (0.9, 0.99):
    acc = 0
    for i, v in enumerate(hot):
        acc += v
        if acc >= frac * total:
            print(f"  {100 * frac:.0f} % of the executed warp instructions come from {i + 1} SASS instructions ({(i + 1) * 16 / 1024:.1f} KB)")
            break

print(f"\nby call site (outermost {depth} frames): executed SASS instructions, share of the executed warp instructions")
stat, dyn = collections.Counter(), collections.Counter()
for a, (n, _) in executed.items():
    if n > 0 and a in chain_of:
        key = " <- ".join(reversed(chain_of[a][:depth])) if not sub_of[a] else f"[{sub_of[a]}]"
        stat[key] += 1
        dyn[key] += n
for key in sorted(stat, key=lambda k: -(stat[k] * 16 / 1024 + 100.0 * dyn[k] / total))[:28]:  # big or hot
    print(f"  {stat[key]:5d} SASS  {100 * dyn[key] / total:5.1f} %  {key}")

print("\nIEEE slow-path subroutines (calls = executions of the subroutine's first instruction)")
subs = collections.defaultdict(list)
for a in sorted(sub_of):
    if sub_of[a]:
        subs[sub_of[a]].append(a)
for name, addrs in subs.items():
    n = sum(executed.get(a, (0, 0))[0] for a in addrs)
    first = executed.get(addrs[0], (0, 0))
    print(f"  {name:40s} {len(addrs):4d} SASS  {first[0]:9d} warp calls  {first[1]:10d} thread calls  {100 * n / total:5.2f} % of the warp instructions")
calls = []
for a, text in text_of.items():
    if "CALL" in text and "slowpath" in text and executed.get(a, (0, 0))[0] > 0:
        calls.append((executed[a][0], executed[a][1], re.sub(r".*\$__cuda_", "", text).rstrip("`) "), " <- ".join(reversed(chain_of[a][:3]))))
for n, t, what, where in sorted(calls, reverse=True)[:10]:
    print(f"    {n:9d} warp calls {t:10d} threads  {what:32s} from {where}")
