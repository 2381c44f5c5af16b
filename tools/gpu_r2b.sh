#!/bin/bash
# Round-2 pass B: whole GPU suite, variant A/B, ncu full capture of the three sweeps (source-level), ncu launch list.
mkdir -p gpurun_out
L=gpurun_out/r2b.log
: > $L
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -25 gpurun_out/pytest_r2b.log | tee -a $L
run() { echo "== $*" | tee -a $L; env "$@" timeout 300 python tools/sweep_bench.py --steps 8 2>&1 | tail -2 | tee -a $L; }
run X=default
for v in t128b3 t96b4 early tabs0 t160; do run MFC5_LIB=$PWD/mfc_b200/libmfc5_$v.so; done
run MFC5_CHUNK=41
run MFC5_CHUNK=51
run MFC5_CHUNK=67
run X=default
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_x|k_march2' -s 6 -c 3 -o gpurun_out/prof_r2b -f python tools/sweep_bench.py --steps 2 > gpurun_out/ncu_r2b.log 2>&1
tail -3 gpurun_out/ncu_r2b.log | tee -a $L
ls -la gpurun_out | tail -5
