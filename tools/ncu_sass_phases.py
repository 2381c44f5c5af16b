#!/usr/bin/env python
"""Per-address-range totals from the SASS source page of an .ncu-rep: executed warp instructions, FP64-pipe instructions and
warp-stall samples per phase of a kernel.
usage: python tools/ncu_sass_phases.py rep.ncu-rep <kernel regex> [<lo-hi>:<name> ...]   (offsets in hex from the kernel start)
With no ranges: one line per backward branch target region (loop structure) is printed to help choose them."""
import csv, io, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(blk)))
print(rows[0][1][:90])
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = []
base = None
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    ins.append((a - base, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
FP = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")
def op(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0]
tot_ex = sum(i[3] for i in ins); tot_s = sum(i[2] for i in ins)
ranges = []
for a in sys.argv[3:]:
    rng, name = a.split(":")
    lo, hi = (int(x, 16) for x in rng.split("-"))
    ranges.append((lo, hi, name))
if not ranges:
    for off, s, smp, ex in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            print(f"  {off:#06x} {s}   {'<- backward' if tgt < off else ''}")
    sys.exit(0)
print(f"{'phase':28s} {'warp inst':>12s} {'share':>6s} {'fp64-pipe':>12s} {'other':>12s} {'samples':>8s} {'share':>6s}  samples/inst")
for lo, hi, name in ranges:
    sel = [i for i in ins if lo <= i[0] <= hi]
    ex = sum(i[3] for i in sel); fp = sum(i[3] for i in sel if op(i[1]) in FP); sm = sum(i[2] for i in sel)
    print(f"{name:28s} {ex:12d} {ex / tot_ex:6.3f} {fp:12d} {ex - fp:12d} {sm:8d} {sm / max(tot_s, 1):6.3f}  {sm / max(ex, 1) * 1e4:8.2f}e-4")
print(f"{'total':28s} {tot_ex:12d} {'':6s} {sum(i[3] for i in ins if op(i[1]) in FP):12d} {'':12s} {tot_s:8d}")
