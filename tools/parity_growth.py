#!/usr/bin/env python
"""Error growth of the product build against the oracle on a reduced C5: python tools/parity_growth.py [m] [steps]
Modes: ref_order (reference operation order, FMA contraction on) and fast (division-light arithmetic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mfc_b200 import cases, pre_process
from mfc_b200.lib import Mfc5
from mfc_b200.params import params_from_case
from oracle_lib import Oracle
m = int(sys.argv[1]) if len(sys.argv) > 1 else 39
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
case = cases.perf_test_3d(m=m, steps=nsteps)
cbs, ccs = pre_process.generate_grid(case)
_, cons = pre_process.initial_condition(case, ccs)
dt = float(case["dt"])
P0 = params_from_case(case, fast_math=0)
o = Oracle(P0, cbs, omp=True); o.set_cons(cons)
gs = {"ref_order": Mfc5(params_from_case(case, fast_math=0), cbs), "fast": Mfc5(params_from_case(case, fast_math=1), cbs)}
for g in gs.values(): g.set_cons(cons)
for it in range(1, nsteps + 1):
    o.step(dt)
    for g in gs.values(): g.steps(dt, 1)
    if it in (1, 2, 5, 10, 20, 50, 100, 200) or it == nsteps:
        b = o.get_cons()
        line = f"step {it:4d} max|q| " + " ".join(f"{np.max(np.abs(b[v])):.1e}" for v in range(b.shape[0]))
        for name, g in gs.items():
            g.sync(); a = g.get_cons()
            errs = [np.max(np.abs(a[v] - b[v])) / (np.max(np.abs(b[v])) + 1e-300) for v in range(a.shape[0])]
            line += f" | {name}: " + " ".join(f"{e:.1e}" for e in errs)
        print(line, flush=True)
