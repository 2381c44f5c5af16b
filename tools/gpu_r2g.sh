#!/bin/bash
# Round-2 pass G: GPU suite on the tree with DevP::fast_riem (fast WENO + reference-order Riemann solve for wave_speeds = 2 / low_Mach) + their timings.
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2g.log
tail -25 gpurun_out/pytest_r2g.log | tee -a $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -2 | tee -a $L; }
run X=default timeout 300 python tools/sweep_bench.py --steps 8
run X=ws2 timeout 300 python tools/sweep_bench.py --steps 4 --set wave_speeds=2
run X=lowmach1 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=1
run X=lowmach2 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=2
run MFC5_NO_FAST_WENO_ONLY=1 timeout 300 python tools/sweep_bench.py --steps 4 --set wave_speeds=2
run MFC5_NO_OPT=1 timeout 300 python tools/sweep_bench.py --steps 4
