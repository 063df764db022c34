"""Executed warp instructions per SOURCE LINE of a kernel: joins the SASS view of an ncu report (per-instruction
"Instructions Executed") with nvdisasm's line table of the object file (-lineinfo).  Tells where an issue-bound kernel
spends its instruction slots.

usage: python tools/ncu_lines.py report.ncu-rep object.o mangled_kernel_name [N]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, fun = sys.argv[1:4]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
base = int(data[0]["Address"], 16)
execd = {int(d["Address"], 16) - base: int(d["Instructions Executed"] or 0) for d in data}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + fun + ":"))
per_line = collections.Counter()
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m:
        per_line[cur] += execd.get(int(m.group(1), 16), 0)
tot = sum(per_line.values())
print("total warp instructions", tot, "(ncu:", sum(execd.values()), ")")
src = {}
for (f, ln), c in per_line.most_common(N):
    if f not in src:
        for root, _, files in os.walk(os.path.dirname(os.path.abspath(obj)) + "/../csrc"):
            if f in files:
                src[f] = open(os.path.join(root, f)).read().splitlines()
    text = src.get(f, [""] * (ln + 1))[ln - 1].strip() if f in src and ln <= len(src[f]) else ""
    print(f"{c:10d} {100 * c / tot:5.1f}%  {f}:{ln:<4d} {text[:120]}")
