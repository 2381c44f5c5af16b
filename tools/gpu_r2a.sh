#!/bin/bash
# Round-2 pass A: GPU suite on the new kernels, then A/B of the marching-kernel variants (per-kernel CUDA-event times).
mkdir -p gpurun_out
L=gpurun_out/r2a.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -15 gpurun_out/pytest_r2a.log | tee -a $L
run() { echo "== $*" | tee -a $L; env "$@" timeout 300 python tools/sweep_bench.py --steps 8 2>&1 | tail -2 | tee -a $L; }
run X=default
run MFC5_MARCH_V1=1
run MFC5_NO_FOLD=1
run MFC5_NO_FUSE_C2P=1
run MFC5_NO_FOLD=1 MFC5_NO_FUSE_C2P=1
for v in n256t0 n224early n192; do run MFC5_LIB=$PWD/mfc_b200/libmfc5_$v.so; done
run X=default
