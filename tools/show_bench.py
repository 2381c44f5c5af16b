#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_n1.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "| GPUs", d["n_gpus"])
print(f"  ms/step {d['ms_per_step']:.3f}  grind {d['value']:.5f} {d['unit']}  launches {d.get('gpu_launches')}")
for k in ("roofline", "roofline_stage", "roofline_fp64"):
    if d.get(k):
        r = d[k]
        print(f"  {k}: frac {r['frac']:.3f}", {x: (round(v, 3) if isinstance(v, float) else v) for x, v in r.items()
                                              if x in ("kernel", "achieved", "peak", "unit", "stage_floor_ms", "per_sweep_frac")})
if d.get("e2e"):
    print(f"  e2e: {d['e2e']['ms_per_step']:.2f} ms/step, grind {d['e2e']['value']:.5f}")
if d.get("cpu_baseline"):
    print("  cpu_baseline:", d["cpu_baseline"])
print("  kernels:", {k: round(v["ms_per_launch"], 4) for k, v in d.get("kernels", {}).items()})
print("  clocks:", d.get("clocks"))
