#!/bin/bash
# Round-2 final one-GPU pass: whole GPU suite incl. the full-size acceptance cases, bench (native + CPU arm), ncu launch list, ncu full capture.
mkdir -p gpurun_out
L=gpurun_out/r2final.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $L
rm -f gpurun_out/acceptance.log
MFC5_ACCEPTANCE=1 timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2final.log
tail -8 gpurun_out/pytest_r2final.log | tee -a $L
cat gpurun_out/acceptance.log | tee -a $L
timeout 600 python bench.py > gpurun_out/bench_r2final.json 2> gpurun_out/bench_r2final.err; tail -c 400 gpurun_out/bench_r2final.json | tee -a $L
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2final_ref.json 2> gpurun_out/bench_r2final_ref.err; tail -c 300 gpurun_out/bench_r2final_ref.json | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r2final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/ncu_l_r2final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_x|k_march2|k_cbc' -s 8 -c 5 -o gpurun_out/prof_r2final -f python tools/sweep_bench.py --steps 2 > gpurun_out/ncu_r2final.log 2>&1
tail -2 gpurun_out/ncu_r2final.log | tee -a $L
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $L
