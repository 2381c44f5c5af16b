#!/bin/bash
# Round-2 pass H (8 GPUs): multi-rank parity tests on the final kernels (worlds 2 / 4 / 8), host<->device bandwidth probe,
# weak-scaling bench at 8 and 2 ranks (bench.py runs its decomposed-vs-single parity check first).
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
nvidia-smi -L | wc -l | tee -a $L
nvidia-smi topo -m > gpurun_out/topo_r2h.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_driver.py::test_two_ranks_match_golden -m gpu -q -rs > gpurun_out/pytest_multi_r2h.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_r2h.log; tail -6 gpurun_out/pytest_multi_r2h.log | tee -a $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py > gpurun_out/pcie_probe_n8.json 2> gpurun_out/pcie_probe_n8.err; tail -c 2500 gpurun_out/pcie_probe_n8.json | tee -a $L
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-other-configs > gpurun_out/bench_r2h_n$n.json 2> gpurun_out/bench_r2h_n$n.err; echo "bench $n rc=$?" | tee -a $L
tail -c 1800 gpurun_out/bench_r2h_n$n.json | tee -a $L
done
MFC5_NO_BIND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-other-configs > gpurun_out/bench_r2h_n8_nobind.json 2> gpurun_out/bench_r2h_n8_nobind.err
tail -c 900 gpurun_out/bench_r2h_n8_nobind.json | tee -a $L
