#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c.log
: > $L
run() { echo "== $*" | tee -a $L; env "$@" timeout 300 python tools/sweep_bench.py --steps 8 2>&1 | tail -1 | tee -a $L; }
run X=default
for v in fps late2 fpslate1 xnv4 xrpb8 xminb3 nv4t128 nv4; do run MFC5_LIB=$PWD/mfc_b200/libmfc5_$v.so; done
run X=default
