#!/bin/bash
# A/B of kernel-variant builds (make -C mfc_b200/csrc variant V=...): parity tests on the default build, then per-kernel times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/variants.log
python tools/sweep_bench.py --steps 10 2>&1 | tee -a gpurun_out/variants.log
for v in "$@"; do
  MFC5_LIB=$PWD/mfc_b200/libmfc5_$v.so python tools/sweep_bench.py --steps 10 2>&1 | tee -a gpurun_out/variants.log
done
python tools/sweep_bench.py --steps 10 2>&1 | tee -a gpurun_out/variants.log
