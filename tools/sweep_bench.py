"""Quick per-kernel timing of the bench workload (no tests, no CPU arm): python tools/sweep_bench.py [m] [case]
Honours MFC5_LIB (alternative build of libmfc5.so) and the library's MFC5_* environment knobs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import bench
from mfc_b200.lib import Mfc5
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=200)
ap.add_argument("--case", default="3D_performance_test")
ap.add_argument("--riemann", default="hllc")
ap.add_argument("--gbpp", type=int, default=16)
ap.add_argument("--weno5z", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE", help="override a case parameter (e.g. wave_speeds=2, low_Mach=1)")
a = ap.parse_args()
args = argparse.Namespace(case=a.case, m=a.m, riemann=a.riemann, gbpp=a.gbpp, weno5z=a.weno5z, scaling="weak")
case, dims, name = bench.build_case(args, 1)
for kv in a.set:
    k, v = kv.split("=", 1)
    try:
        v = int(v)
    except ValueError:
        try:
            v = float(v)
        except ValueError:
            pass
    case[k] = v
    name += f" [{k}={v}]"
blk, lcase, P, cbs, gws, cons = bench.setup_rank(case, 0, 1, dims)
g = Mfc5(P, cbs)
g.set_cons(cons)
dt = float(case["dt"])
g.steps(dt, 2); g.sync()
tot = {}
for _ in range(a.steps):
    g.step(dt)
    for k, (ms, n) in g.last_kernel_ms().items():
        t = tot.setdefault(k, [0.0, 0]); t[0] += ms; t[1] += n
stage = sum(v[0] for v in tot.values()) / a.steps / P.time_stepper
ncell = (P.m + 1) * (P.n + 1) * (P.p + 1)
print(os.environ.get("MFC5_LIB", "libmfc5.so"), name)
print("  " + "  ".join(f"{k}={v[0] / max(v[1], 1):.3f}ms" for k, v in tot.items() if v[1]), f" stage={stage:.3f}ms grind={stage * 1e6 / (g.N * P.m * max(P.n, 1) * max(P.p, 1)):.4f} ns  nan={g.reduce(2)}")
