#!/usr/bin/env python
"""Opcode mix of an address range of one kernel: python tools/sass_loop.py file.sass <name substring> <lo hex> <hi hex> [-v]"""
import re, sys, collections
path, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
lines = open(path).read().split("\n")
idx = [i for i, l in enumerate(lines) if "Function :" in l]
start = [i for i in idx if key in lines[i]][0]
nxt = [i for i in idx if i > start]
end = nxt[0] if nxt else len(lines)
ins = []
for l in lines[start:end]:
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
loop = [t for a, t in ins if lo <= a <= hi]
h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in loop)
print("instructions", len(loop))
print("  ".join(f"{k}:{v}" for k, v in h.most_common()))
if "-v" in sys.argv:
    for a, t in ins:
        if lo <= a <= hi: print(hex(a), t)
