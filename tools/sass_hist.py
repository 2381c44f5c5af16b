#!/usr/bin/env python
"""Opcode histogram of one kernel in a cuobjdump -sass listing: python tools/sass_hist.py file.sass <substring of mangled name> [lo hi]
(lo/hi: optional line range inside the function, to isolate a loop body)"""
import re, sys, collections
path, key = sys.argv[1], sys.argv[2]
lines = open(path).read().split("\n")
start = None
for i, l in enumerate(lines):
    if "Function :" in l:
        if start is not None: end = i; break
        if key in l: start = i
else:
    end = len(lines)
body = lines[start:end]
if len(sys.argv) > 4: body = body[int(sys.argv[3]):int(sys.argv[4])]
h = collections.Counter()
for l in body:
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m: h[m.group(2).split(".")[0]] += 1
tot = sum(h.values())
fp64 = sum(v for k, v in h.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "DMNMX"))
print(f"total {tot}  fp64-pipe {fp64}")
for k, v in h.most_common(): print(f"  {k:10s} {v}")
