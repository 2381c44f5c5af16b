#!/usr/bin/env python3
"""Regenerate tests/golden/ fixtures from the read-only reference tree (run in the build container only).

What it does (nothing here is copied reference *source*; goldens are test data):
  1. expands the reference's own test-case generator (toolchain/mfc/test/cases.py:list_cases, with
     the optional third-party modules it imports stubbed out) to obtain the
     full parameter dictionary of every case whose UUID is inside the hot-path envelope
     (SURVEY.md section 4), and writes them to cases.json;
  2. converts tests/<UUID>/golden.txt (one line per D/*.dat file: name + values,
     toolchain/mfc/packer/pack.py:117-158) into a compressed <UUID>.npz of float64 arrays.

Usage:  python tests/golden/make_fixtures.py [/root/reference]
"""
import json
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

# UUIDs inside the 5-eq / WENO{3,5}{JS,M,Z} / HLL|HLLC / RK3 envelope;
# (int_comp = 2 (MTHINC): the CPU oracle only so far, tests/test_oracle_next.py)
# last line: viscous (finite-difference velocity gradients, weno_Re_flux = F)
# before them: the f-4 widening (low_Mach = 1 / 2, tolerance 1e-10 by the reference's own rule; alt_soundspeed) and f-3 (elliptic smoothing, hard-coded profiles), Lax-Friedrichs, MUSCL
UUIDS = """
D79C3E6F 8A59E8E6 2AB32975 70DAE9E8 3AE495F4 5EC236F2 48CCE072 AED93D34 C5B79059 727F72ED A60691E7 3FC6FC4A
2FD933A2 ED728400 9D92767D 63CD6ADF
55533234 EAA53889 D972BA0F 20AE0551 A6E65782 4129A23A E84967E7 5F877BC9 16C03D8E B96AC58F 8FDEE23A BF46F657
4D7926CD 98081DA3 10DBA76B 212401B4 EB6687AB
6FC6A809 09DAFEBA 6B4B738B F99FBB36 E09A12D9 5010B814 730DFD6D ABAC3AE3 C93BE9B5 D0045756 557FF170 61FFF3D3
6249EDDE 8BCBAED8 9DCBE54F 0BDC8E40 38C2159F D7107B89
2A5CFF02 F9850EA5 16A922E4 7789B55A 9090FC4B 677DAA82
EA9C3FA3 A21CF713 39816683 9EF5305B 31742FC7 D6415F48
4B7FC359 BB0418CD EFDCFF72 35231DF8 5F099890 04E1CB26
5527832F 4AEF478A 32D0F235 18BDCBC8 F97573DB F4F6AC27 1E738705 0879E062 83EFC30C 1CCA82F5 3A8359F6 6D24B115
461DCB09 C4907722
7374E266 3BFEAC19 FBF808BE 3B414AF0 3C00B89D 345A94C0 16FBF4C8 DC9CB97E A5C93D62 A6AC2E06 5781A4C2 645A26E3
FC4D07B6 F0F175B2 AE3ECF01
9ACD5174 73B0539E 2A523AC1 C06849AD AB0BE4E4 C36F18FB C4A2FAA3 851F7AE2 BD8004FF 758D0268 AACF1BC5 B33E256A
B8F5F1C8 A0B82851 C7B686C0
0FCCE9F1 8C7AA13B CE232828
05D092C5 59C3F366 B49877F3 1A379909 C3A4DBAE
BCA926AC 24FD489A B5F0C655 82152069 BCB15E8E 02D1E522
4D88A48D 07C1FECC 0853FBD1 F9E0BE95 F5F06EA3 C462AA3E
7C8F1BA9 A26B7E00 E47F81A2
8C1E22E4 C1134EF2 9843489C B50DE910 1DBC8382 EF3E3FF1
C9E0B721 A078904B 8370A3D9 8FD91C41
AFB46CF8 4AADC848
F1D01264 14846B52 90D62933 A7C19B7B 4F79391A 77292B06 9BD657CB 30035426 01DC0251 24119641 F33CC4CA
7BCDC8E7 10041BB7 B54BB9D8
516A2067 A089CE54 10DE58AA
48D5C130 E719C5F2 EDADA16E
B3E70A3A 0ACB1F16 A3EC9BD7
7B71820F 56BB8D80 0347811A 5D68B0F5 A72F7006 B8BF00C9 9E9DA539 448EA9A4 4A15969C E3EB6065 CD6C83A2 0F6DB678 6247B427
5126B21F 4E4FECA9 4F3722DB 4C4F339C 7A1719C6
3F0FA534 D78D6B12 6FA80DE9 3EE4302C F4E54FC0 F207F0FE B46CEFF3 28D9A304 71258175
F883F4EE 743D6D71 84A12B88 6B227FF5 D68EDC50 5EC347F2 42848784 374C9B16 347964C3
AFE4994E 66CB22E3 7B21681C CB76A5E1 96AF00B3 9DAC4DDC F5512823 F28381D9 0288BDAD
09DDCC0C 1A78D4CC 126FF74F 290C3BA8 9730B2DF 5DAB50B2 4912EFF1 B2AEF240 1929BEC7 55A51767
EBBC93BD F1CF01C4 DC97B66D 4111B2ED
4201674B 342686E8 DEA0A0D6 AE02324F E21940B3 121D4ECA
8EAC3DA7 C79E1D3C CD6DC908 9CB03CEF 0A022883 1C0780C8 70EC99CE
""".split()


def expand_cases():
    sys.dont_write_bytecode = True
    for name in ["fastjsonschema", "mako", "mako.template", "mako.lookup", "cantera", "pyrometheus"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["fastjsonschema"].compile = lambda *a, **k: (lambda x: x)
    sys.modules["fastjsonschema"].JsonSchemaException = Exception
    sys.modules["mako.template"].Template = object
    sys.modules["mako.lookup"].TemplateLookup = object
    sys.path.insert(0, os.path.join(REF, "toolchain"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from mfc.test.cases import list_cases  # noqa

        out = {}
        for b in list_cases():
            try:
                c = b.to_case()
            except Exception:  # example cases needing cantera etc. -- none in scope
                continue
            u = c.get_uuid()
            if u in UUIDS:
                tol = c.override_tol
                if tol is None:
                    tol = 1e-3 if "Example" in c.trace else 1e-12
                out[u] = dict(trace=c.trace, ppn=c.ppn, tol=tol, params=c.params)
    finally:
        os.chdir(cwd)
    return out


def convert_golden(uuid):
    arrs = {}
    with open(os.path.join(REF, "tests", uuid, "golden.txt")) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            key = re.sub(r"^D/", "", parts[0]).replace(".dat", "")
            arrs[key] = np.array([float(x) for x in parts[1:]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, uuid + ".npz"), **arrs)
    return len(arrs)


if __name__ == "__main__":
    cases = expand_cases()
    missing = [u for u in UUIDS if u not in cases]
    assert not missing, missing
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)
    for u in UUIDS:
        n = convert_golden(u)
        print(u, cases[u]["trace"], n)
