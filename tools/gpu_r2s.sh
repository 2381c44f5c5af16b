#!/bin/bash
# Round-2 pass S: ncu launch list + full capture on the final binary (k_cbc at 3 blocks per SM), smoke.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r2s.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/ncu_l_r2s.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_x|k_march2|k_cbc' -s 8 -c 5 -o gpurun_out/prof_r2s -f python tools/sweep_bench.py --steps 2 > gpurun_out/ncu_r2s.log 2>&1
tail -2 gpurun_out/ncu_r2s.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
