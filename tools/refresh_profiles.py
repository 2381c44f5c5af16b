#!/usr/bin/env python
"""Turn the scratch output of tools/gpu_round.sh (gpurun_out/) into the tracked round artefacts under profiles/:
launch list (md + csv), ncu sweep summary, per-kernel DRAM traffic (traffic.json), FP64 instruction counts, bench lines.
usage: python tools/refresh_profiles.py r02 [bench.json launches.csv report.ncu-rep [reference.json]]   (paths under gpurun_out/)"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
py = sys.executable
f_bench, f_launch, f_rep, f_ref = (sys.argv[2:6] + [None] * 4)[:4]
f_bench, f_launch, f_rep, f_ref = f_bench or "bench_n1.json", f_launch or "launches.csv", f_rep or "prof_full.ncu-rep", f_ref or "bench_ref.json"
shutil.copy(os.path.join(G, f_bench), os.path.join(P, f"bench_{tag}_n1.json"))
if os.path.exists(os.path.join(G, f_ref)):
    shutil.copy(os.path.join(G, f_ref), os.path.join(P, f"bench_{tag}_reference.json"))
shutil.copy(os.path.join(G, f_launch), os.path.join(P, f"launches_{tag}.csv"))
d = json.loads(open(os.path.join(G, f_bench)).read().strip().splitlines()[-1])
ll = subprocess.run([py, os.path.join(ROOT, "tools", "launch_list.py"), os.path.join(G, f_launch)], capture_output=True, text=True).stdout
tot = sum(v["ms_per_step"] for v in d["kernels"].values())
with open(os.path.join(P, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list, {tag} (`ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e`)\n\n")
    f.write("Workload: examples/3D_performance_test, 201^3 cells, WENO5-M + HLLC, one B200. Times are cold-cache and serialised under ncu: compare SHARES "
            f"with the `kernels` block of `profiles/bench_{tag}_n1.json` (CUDA events, no profiler).\n\n" + ll + "\n")
    f.write("bench.py without the profiler (same box, same call):\n\n| kernel class | ms per launch | launches per step | share of step |\n|---|---|---|---|\n")
    for k, v in d["kernels"].items():
        f.write(f"| {k} | {v['ms_per_launch']:.4f} | {v['launches_per_step']:.0f} | {v['ms_per_step'] / tot:.3f} |\n")
rep = os.path.join(G, f_rep)
summ = subprocess.run([py, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, f"ncu_sweeps_{tag}.md"), "w").write(summ)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    return float(r[ix[k]].replace(",", ""))
units = rows[1]
def dram_bytes(r, k):
    u = units[ix[k]].lower()
    mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return val(r, k) * mult
names = {"k_cons2prim": "cons2prim", "k_sweep_x": "sweep_x", "k_sweep_march<2, 3, 1,": "sweep_y", "k_sweep_march<2, 3, 2,": "sweep_z",
         "k_march2<2, 3, 1,": "sweep_y", "k_march2<2, 3, 2,": "sweep_z"}
traffic, counts = {}, {}
ncell = d["config"]["cells_per_gpu"]
for r in data:
    kn = r[ix["Kernel Name"]]
    for pat, nm in names.items():
        if pat in kn:
            traffic[nm] = dram_bytes(r, "dram__bytes_read.sum") + dram_bytes(r, "dram__bytes_write.sum")
            # thread instructions = (per-cycle rate summed over SMSPs) x elapsed SM cycles (duration x SM clock)
            cyc = val(r, "gpu__time_duration.sum") * 1e-9 * val(r, "sm__cycles_elapsed.avg.per_second") * \
                {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0) * \
                {"hz": 1.0, "khz": 1e3, "mhz": 1e6, "ghz": 1e9}.get(units[ix["sm__cycles_elapsed.avg.per_second"]].lower(), 1.0)
            tot_i = 0.0
            for op in ("dfma", "dmul", "dadd"):
                k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
                if k in ix:
                    tot_i += val(r, k) * cyc
            if tot_i and nm != "cons2prim":
                counts[nm] = round(tot_i / ncell)
traffic["_source"] = f"committed ncu --set full capture {f_rep} ({tag}), dram__bytes_read.sum + dram__bytes_write.sum per launch"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
if counts:
    c = json.load(open(os.path.join(P, "fp64_counts.json")))
    c["thread_inst_per_cell"] = counts
    c["source"] = c["source"].replace("ncu_sweeps_r01", f"ncu_sweeps_{tag}")
    json.dump(c, open(os.path.join(P, "fp64_counts.json"), "w"), indent=1)
print("traffic", traffic)
print("fp64 thread instructions per cell", counts)
