#!/bin/bash
# Strong scaling of C4 (examples/3D_shockdroplet, 201x101x101 cells in total) on 8 / 4 / 2 GPUs of one box.
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --case 3D_shockdroplet --scaling strong --steps 100 --warmup 5 --no-e2e > gpurun_out/bench_c4_strong_n$n.json 2> gpurun_out/bench_c4_strong_n$n.err; echo "n=$n rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_strong_n$n.json").read().strip().splitlines()[-1])
    print("  ", d["config"]["proc_grid"], "ms/step", round(d["ms_per_step"],4), "grind", round(d["value"],5), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
except Exception as e: print("   failed", e)
PY
done
