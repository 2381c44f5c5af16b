#!/bin/bash
# Round-2 pass R: k_cbc at 3 blocks per SM -- parity suite + bench line.
mkdir -p gpurun_out
L=gpurun_out/r2r.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_three_fluids.py -m gpu -q > gpurun_out/pytest_r2r.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2r.log; tail -3 gpurun_out/pytest_r2r.log | tee -a $L
timeout 600 python bench.py > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err
python tools/show_bench.py gpurun_out/bench_r2r.json | head -12 | tee -a $L
