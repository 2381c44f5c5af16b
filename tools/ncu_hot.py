#!/usr/bin/env python
"""Per-instruction stall hot spots of one kernel in an .ncu-rep.  usage: ncu_hot.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
import re
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = next(i for i in secs if re.search(rx, rows[i][1]))
end = next((j for j in secs if j > sel), len(rows))
print(rows[sel][1])
hdr = rows[sel + 1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[sel + 2:end] if r and r[0].startswith("0x")]
S = lambda r, h: int(r[idx[h]] or 0)
tot = sum(S(r, "# Samples") for r in data)
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((h, sum(S(r, h) for r in data)) for h in cols), key=lambda x: -x[1])
print("instructions", len(data), "samples", tot)
print("stall totals:", [(h, v, f"{100*v/tot:.1f}%") for h, v in agg[:9]])
top = sorted(range(len(data)), key=lambda i: -S(data[i], "# Samples"))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((h, S(r, h)) for h in cols), key=lambda x: -x[1])[:2]
    print(f"{i:5d} {r[idx['Source']][:64]:64s} {S(r,'# Samples'):6d} {st}")
