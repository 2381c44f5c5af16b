#!/bin/bash
# The other BASELINE configurations on one GPU: C3 (benchmark, reference size rule for 180 GB), C4 (shock droplet as shipped and
# with the WENO5-Z override), C5 with HLL.  No CPU arm (bench.py's default run has it).
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("  ", d["config"]["workload"], "| ms/step", round(d["ms_per_step"],3), "| grind", round(d["value"],5), "| stage frac", round(d["roofline_stage"]["frac"],3), "| e2e ms", round(d["e2e"]["ms_per_step"],2) if "e2e" in d else None)
    print("  ", {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
except Exception as e: print("   failed", e)
PY
}
run c5_hll --riemann hll
run c4_weno3 --case 3D_shockdroplet --steps 50
run c4_weno5z --case 3D_shockdroplet --weno5z --steps 50
run c3_gbpp180 --case 5eq_rk3_weno3_hllc --gbpp 180 --steps 5 --warmup 3 --no-e2e
