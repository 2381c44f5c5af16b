#!/usr/bin/env python
"""Warp-sampling share of address regions of one kernel in an .ncu-rep.
usage: ncu_regions.py rep kernel_regex lo1:hi1[:label] lo2:hi2 ...   (hex SASS addresses, inclusive)"""
import csv, io, re, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
regs = []
for a in sys.argv[3:]:
    p = a.split(":")
    regs.append((int(p[0], 16), int(p[1], 16), p[2] if len(p) > 2 else a))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = next(i for i in secs if re.search(rx, rows[i][1]))
end = next((j for j in secs if j > sel), len(rows))
hdr = rows[sel + 1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[sel + 2:end] if r and r[0].startswith("0x")]
S = lambda r, h: int(r[idx[h]] or 0)
tot = sum(S(r, "# Samples") for r in data)
exe = "Instructions Executed" if "Instructions Executed" in idx else None
base = int(data[0][0], 16)
for lo, hi, lab in regs:
    sub = [r for r in data if lo <= int(r[0], 16) - base <= hi]
    s = sum(S(r, "# Samples") for r in sub)
    e = sum(S(r, exe) for r in sub) if exe else 0
    fp = sum(S(r, exe) for r in sub if re.match(r"\s*(@!?U?P\d+\s+)?(DFMA|DMUL|DADD|DSETP|MUFU)", r[idx["Source"]])) if exe else 0
    print(f"{lab:24s} static {len(sub):5d}  samples {s:7d} ({100.0 * s / tot:5.1f}%)  executed {e:12d}  fp64-pipe {fp:12d}")
print("total samples", tot)
