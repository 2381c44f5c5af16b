#!/bin/bash
# Round-2 pass P: the benchmark case (C3, benchmarks/5eq_rk3_weno3_hllc) at a "fills-HBM" size (the reference's memory rule stops
# at 90 M cells on 180 GB; this library needs 4 state buffers = 256 B per cell).  Size chosen from the box's free host memory (the
# initial condition is generated on the host).
mkdir -p gpurun_out
L=gpurun_out/r2p.log
: > $L
free -g | tee -a $L
nvidia-smi --query-gpu=memory.total,memory.used --format=csv | tee -a $L
AV=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$AV" -ge 400 ]; then G=900; elif [ "$AV" -ge 150 ]; then G=480; else G=0; fi
echo "MemAvailable ${AV} GB -> gbpp $G" | tee -a $L
if [ "$G" -gt 0 ]; then
  ( timeout 900 python bench.py --case 5eq_rk3_weno3_hllc --gbpp $G --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/bench_r2p_c3_big.json 2> gpurun_out/bench_r2p_c3_big.err ); echo "rc=$?" | tee -a $L
  tail -3 gpurun_out/bench_r2p_c3_big.err | tee -a $L
  python tools/show_bench.py gpurun_out/bench_r2p_c3_big.json | head -6 | tee -a $L
  nvidia-smi --query-gpu=memory.total,memory.used --format=csv | tee -a $L
fi
