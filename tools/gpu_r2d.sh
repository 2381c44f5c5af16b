#!/bin/bash
# Round-2 pass D: whole GPU suite on the tree with WENO7 in CUDA + bench line.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2d.log
tail -40 gpurun_out/pytest_r2d.log
timeout 600 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 2500 gpurun_out/bench_r2d.json
