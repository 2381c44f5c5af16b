#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -5 gpurun_out/bench_r2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_stage']['frac'])
for k,v in d.get('other_configs',{}).items():
    print(k, {x:v.get(x) for x in ('cells_per_gpu','ms_per_step','value','setup_s','graph_replay','skipped')}, v.get('roofline_stage',{}).get('frac'))
PY
