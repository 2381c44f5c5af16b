#!/bin/bash
# Round-2 pass X: x kernel edge states as one double2 per lane (16-byte shared-memory accesses) -- A/B + parity.
mkdir -p gpurun_out
L=gpurun_out/r2x.log
: > $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -1 | tee -a $L; }
for i in 1 2; do
run X=e2 timeout 200 python tools/sweep_bench.py --steps 10
run MFC5_LIB=$PWD/mfc_b200/libmfc5_xe0.so timeout 200 python tools/sweep_bench.py --steps 10
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_symmetry.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_r2x.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2x.log; tail -3 gpurun_out/pytest_r2x.log | tee -a $L
