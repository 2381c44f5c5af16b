#!/bin/bash
# Round-2 closing pass: the driver's GPU test command, bench line, ncu launch list + full capture on the final binary.
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2w.log 2>&1; echo rc=$? >> gpurun_out/pytest_r2w.log; tail -3 gpurun_out/pytest_r2w.log
timeout 600 python bench.py > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; python tools/show_bench.py gpurun_out/bench_r2w.json | head -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r2w.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/ncu_l_r2w.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_x|k_march2|k_cbc' -s 8 -c 5 -o gpurun_out/prof_r2w -f python tools/sweep_bench.py --steps 2 > gpurun_out/ncu_r2w.log 2>&1
tail -1 gpurun_out/ncu_r2w.log
