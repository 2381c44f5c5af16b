#!/bin/bash
# Round-2 pass F: whole GPU suite (viscous, templated k_cbc), bench line, slow-path numbers, ncu launch list + full capture.
mkdir -p gpurun_out
L=gpurun_out/r2f.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $L
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f.log
tail -25 gpurun_out/pytest_r2f.log | tee -a $L
timeout 600 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 600 gpurun_out/bench_r2f.json | tee -a $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -2 | tee -a $L; }
run X=default timeout 300 python tools/sweep_bench.py --steps 8
run MFC5_NO_FAST=1 timeout 300 python tools/sweep_bench.py --steps 4
run X=ws2 timeout 300 python tools/sweep_bench.py --steps 4 --set wave_speeds=2
run X=lowmach1 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=1
run X=lowmach2 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=2
run X=hll timeout 300 python tools/sweep_bench.py --steps 4 --riemann hll
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/ncu_l_r2f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_x|k_march2|k_cbc' -s 8 -c 5 -o gpurun_out/prof_r2f -f python tools/sweep_bench.py --steps 2 > gpurun_out/ncu_r2f.log 2>&1
tail -3 gpurun_out/ncu_r2f.log | tee -a $L
ls -la gpurun_out | tail -8
