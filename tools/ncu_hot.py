"""Top stall sites of a kernel from an ncu report's source page (SASS view).

usage: python tools/ncu_hot.py report.ncu-rep [N]  -> the N SASS instructions with most stall samples."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"] or 0) for d in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot}")
agg = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, d in sorted(enumerate(data), key=lambda kv: -int(kv[1]["# Samples"] or 0))[:N]:
    top = sorted(((int(d[s] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{int(d['# Samples']):7d} {100 * int(d['# Samples']) / max(tot, 1):5.1f}%  #{i:5d} {d['Source'][:90]:90s} {top}")
