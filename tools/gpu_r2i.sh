#!/bin/bash
# Round-2 pass I (2 GPUs): boundary x segments in one launch, fast-path wave_speeds = 2 / low_Mach -- multi-rank parity, N = 2 bench, timings.
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_driver.py::test_two_ranks_match_golden -m gpu -q -rs > gpurun_out/pytest_multi_r2i.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_r2i.log; tail -6 gpurun_out/pytest_multi_r2i.log | tee -a $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x > gpurun_out/pytest_r2i.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2i.log; tail -4 gpurun_out/pytest_r2i.log | tee -a $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-configs --no-e2e 2> gpurun_out/bench_r2i_n2.err | grep '^{' > gpurun_out/bench_r2i_n2.json; echo "bench 2 rc=$?" | tee -a $L
python tools/show_bench.py gpurun_out/bench_r2i_n2.json | tee -a $L
run() { echo "== $*" | tee -a $L; env "$@" 2>&1 | tail -2 | tee -a $L; }
run X=ws2 timeout 300 python tools/sweep_bench.py --steps 4 --set wave_speeds=2
run X=lowmach1 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=1
run X=lowmach2 timeout 300 python tools/sweep_bench.py --steps 4 --set low_Mach=2
run X=hll_ws2 timeout 300 python tools/sweep_bench.py --steps 4 --riemann hll --set wave_speeds=2
