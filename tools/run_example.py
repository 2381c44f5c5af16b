#!/usr/bin/env python
"""Run one of the BASELINE configurations end to end through the host driver (pre_process + simulation on the GPU),
like `./mfc.sh run examples/<case>/case.py` does for the reference:  python tools/run_example.py [3D_performance_test|...] [--m 200]
Prints the driver's own output (progress lines, `Performance:` line) and the wall-clock split."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfc_b200 import cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case", nargs="?", default="3D_performance_test")
ap.add_argument("--m", type=int, default=200)
ap.add_argument("--steps", type=int, default=None)
a = ap.parse_args()
if a.case == "3D_performance_test":
    case = cases.perf_test_3d(m=a.m)
elif a.case == "3D_shockdroplet":
    case = cases.shockdroplet_3d()
elif a.case == "2D_shockbubble":
    case = cases.shockbubble_2d()
else:
    case = cases.sod_1d()
if a.steps:
    case["t_step_stop"] = a.steps
    case["t_step_save"] = a.steps
d = tempfile.mkdtemp(prefix="mfc_b200_")
with open(os.path.join(d, "case.json"), "w") as f:
    json.dump(case, f)
env = dict(os.environ, PYTHONPATH=ROOT)
for tgt in ("pre_process", "simulation"):
    t = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "mfc_b200", "run", os.path.join(d, "case.json"), "-t", tgt], env=env, cwd=ROOT, capture_output=True, text=True)
    dt = time.perf_counter() - t
    lines = r.stdout.strip().split("\n")
    print(f"== {tgt}: rc={r.returncode} wall={dt:.2f}s")
    print("\n".join(lines[:3] + (["..."] if len(lines) > 6 else []) + lines[-3:]))
    if r.returncode:
        print(r.stderr[-3000:])
        sys.exit(1)
print(open(os.path.join(d, "time_data.dat")).read())
sz = sum(os.path.getsize(os.path.join(p, f)) for p, _, fs in os.walk(d) for f in fs)
print(f"case dir {d}: {sz / 1e6:.1f} MB on disk, t_step_stop={case['t_step_stop']}, parallel_io={case.get('parallel_io')}")
