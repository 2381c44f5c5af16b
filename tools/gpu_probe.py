"""Diagnostic run on the GPU box: CUDA vs oracle on a handful of cases, prints error norms."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from golden_util import CASES, compare, golden, golden_field, setup_case
from mfc_b200 import driver
from mfc_b200.lib import Mfc5
from oracle_lib import Oracle

ids = sys.argv[1:] or ["2AB32975", "677DAA82", "C4907722", "5EC236F2", "D6415F48", "4129A23A", "10DBA76B", "04E1CB26", "A0B82851", "5010B814", "B49877F3", "B5F0C655"]
for u in ids:
    case, P, cbs, cons, dbuf = setup_case(u)
    for strict in (True, False):
        try:
            g = Mfc5(P, cbs, dirichlet=dbuf, strict=strict)
            g.set_cons(cons)
            o = Oracle(P, cbs, omp=True)
            for (d, s), b in dbuf.items(): o.set_dirichlet(d, s, b)
            o.set_cons(cons)
            g.compute_rhs(); o.compute_rhs()
            dp = np.max(np.abs(g.get_prim() - o.get_prim()))
            ro = o.get_rhs(); rg = g.get_rhs()
            dr = np.max(np.abs(rg - ro)) / (np.max(np.abs(ro)) + 1e-300)
            t = time.time(); driver.run(g, case); tg = time.time() - t
            driver.run(o, case)
            a, b = g.get_cons(), o.get_cons()
            gold = golden(u); nstep = int(case["t_step_stop"]); shape = a.shape[1:]
            okg = all(compare(a[v], golden_field(gold, "cons", v + 1, 0, nstep, shape), CASES[u]["tol"])[0] for v in range(a.shape[0]))
            worst = max(compare(a[v], b[v], 1e-10)[2] for v in range(a.shape[0]))
            print(f"{u} strict={int(strict)} {CASES[u]['trace']:45s} dprim={dp:.2e} drhs_rel={dr:.2e} 50step_vs_oracle={worst:.2e} golden={'OK' if okg else 'FAIL'} gpu_s={tg:.2f} launches={g.kernel_launches}", flush=True)
            g.close()
        except Exception as e:
            print(u, "strict", strict, "EXC", repr(e)[:300], flush=True)
