#!/bin/bash
# Round-2 pass M (8 GPUs): multi-rank parity incl. viscous at 2 / 4 / 8 ranks on the final tree, weak-scaling bench at 8 and 4 ranks.
mkdir -p gpurun_out
L=gpurun_out/r2m.log
: > $L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_driver.py::test_two_ranks_match_golden -m gpu -q -rs > gpurun_out/pytest_multi_r2m.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_r2m.log; tail -6 gpurun_out/pytest_multi_r2m.log | tee -a $L
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/bench_r2m_n$n.err | grep '^{' > gpurun_out/bench_r2m_n$n.json; echo "bench $n rc=$?" | tee -a $L
python tools/show_bench.py gpurun_out/bench_r2m_n$n.json | sed -n 1,3p | tee -a $L
done
