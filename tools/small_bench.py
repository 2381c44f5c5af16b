"""Wall-clock per step of the small BASELINE configs (launch-latency bound): eager stream vs CUDA-graph replay."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfc_b200 import cases, pre_process
from mfc_b200.lib import Mfc5
from mfc_b200.params import params_from_case
for name, case in (("C1 1D_sodshocktube", cases.sod_1d()), ("C2 2D_shockbubble", cases.shockbubble_2d()),
                   ("C4 3D_shockdroplet WENO3-M", cases.shockdroplet_3d()),
                   ("C4 3D_shockdroplet WENO5-Z", cases.shockdroplet_3d(weno_order=5, mapped_weno="F", wenoz="T"))):
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs)
    g.set_cons(cons)
    dt = float(case["dt"])
    g.steps(dt, 5); g.sync()
    n = 200
    t = time.perf_counter(); g.steps(dt, n); g.sync(); tg = (time.perf_counter() - t) / n
    ncell = (P.m + 1) * (P.n + 1) * (P.p + 1)
    grind = tg / P.time_stepper * 1e9 / (g.N * P.m * max(P.n, 1) * max(P.p, 1))
    print(f"{name:32s} cells={ncell:8d}  {'graph' if not os.environ.get('MFC5_NO_GRAPH') else 'eager'}: {tg * 1e6:8.1f} us/step  grind={grind:.4f} ns  icfl={g.icfl_max(dt):.3f} nan={g.reduce(2)}")
    g.close()
