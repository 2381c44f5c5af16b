#!/bin/bash
# Multi-GPU pass on one box: multi-rank parity tests, then the weak-scaling bench at 8 and 4 ranks.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -3 gpurun_out/pytest_multi.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench $n rc=$?"
tail -c 1500 gpurun_out/bench_n$n.json
done
