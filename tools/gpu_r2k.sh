#!/bin/bash
# Round-2 pass K (2 GPUs): viscous on processor boundaries (direction-by-direction fill + exchange with edges / corners).
mkdir -p gpurun_out
L=gpurun_out/r2k.log
: > $L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs > gpurun_out/pytest_multi_r2k.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_r2k.log; tail -12 gpurun_out/pytest_multi_r2k.log | tee -a $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -q -k "viscous or 1C0780C8 or 70EC99CE or two_ranks" > gpurun_out/pytest_r2k.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2k.log; tail -4 gpurun_out/pytest_r2k.log | tee -a $L
