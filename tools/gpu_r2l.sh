#!/bin/bash
# Round-2 pass L: A/B of the k_march2 instruction-diet build (running field pointers, carried ghost map) against the default.
mkdir -p gpurun_out
L=gpurun_out/r2l.log
: > $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -1 | tee -a $L; }
for i in 1 2; do
run X=default timeout 300 python tools/sweep_bench.py --steps 10
run MFC5_LIB=$PWD/mfc_b200/libmfc5_diet.so timeout 300 python tools/sweep_bench.py --steps 10
done
run X=default_c4 timeout 300 python tools/sweep_bench.py --steps 10 --case 3D_shockdroplet
run MFC5_LIB=$PWD/mfc_b200/libmfc5_diet.so timeout 300 python tools/sweep_bench.py --steps 10 --case 3D_shockdroplet
MFC5_LIB=$PWD/mfc_b200/libmfc5_diet.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "baseline_config or full_size or 100_rk3 or golden" > gpurun_out/pytest_r2l.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2l.log; tail -3 gpurun_out/pytest_r2l.log | tee -a $L
