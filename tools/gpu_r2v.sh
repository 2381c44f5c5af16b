#!/bin/bash
# Round-2 pass V: x kernel with 16-byte window staging (MFC5_X_VEC) -- A/B against the 8-byte build, then the parity suites.
mkdir -p gpurun_out
L=gpurun_out/r2v.log
: > $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -1 | tee -a $L; }
for i in 1 2; do
run X=vec16 timeout 300 python tools/sweep_bench.py --steps 10
run MFC5_LIB=$PWD/mfc_b200/libmfc5_xvec0.so timeout 300 python tools/sweep_bench.py --steps 10
done
run X=vec16_c4 timeout 300 python tools/sweep_bench.py --steps 10 --case 3D_shockdroplet
run MFC5_LIB=$PWD/mfc_b200/libmfc5_xvec0.so timeout 300 python tools/sweep_bench.py --steps 10 --case 3D_shockdroplet
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_symmetry.py tests/test_three_fluids.py -m gpu -q -x > gpurun_out/pytest_r2v.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2v.log; tail -3 gpurun_out/pytest_r2v.log | tee -a $L
