"""What the box's host <-> device path can deliver, next to bench.py's e2e number (which is bound by it).
    python tools/pcie_probe.py                                        # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
Per rank: pinned H2D and D2H bandwidth of a 520 MB buffer (the 201^3 state), (a) each rank alone, one after the other,
(b) all ranks at once; with and without binding the rank to its GPU's cores (mfc_b200/affinity.py).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
NB = 8 * 8120601 * 8


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(host, dev, reps=4):
    out = {}
    for name, (dst, src) in (("h2d", (dev, host)), ("d2h", (host, dev))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        out[name] = NB * reps / (time.perf_counter() - t) / 1e9
    return out


def gather(x):
    if world == 1:
        return [x]
    box = [None] * world
    dist.all_gather_object(box, x)
    return box


res = {"world": world, "bytes": NB}
dev = torch.empty(NB // 8, dtype=torch.float64, device="cuda")
for tag in ("unbound", "bound"):
    if tag == "bound":
        from mfc_b200.affinity import bind_to_gpu
        res["affinity"] = gather(bind_to_gpu(lr, world))
    host = torch.empty(NB // 8, dtype=torch.float64, pin_memory=True)
    host.zero_()
    alone = None
    for r in range(world):   # one rank at a time
        barrier()
        if r == rank:
            alone = measure(host, dev)
    barrier()
    together = measure(host, dev)
    barrier()
    res[tag] = {"alone": gather(alone), "together": gather(together)}
    del host
if rank == 0:
    for tag in ("unbound", "bound"):
        for k in ("alone", "together"):
            rows = res[tag][k]
            res[tag][k] = {"h2d_gbs": [round(x["h2d"], 1) for x in rows], "d2h_gbs": [round(x["d2h"], 1) for x in rows],
                           "h2d_sum": round(sum(x["h2d"] for x in rows), 1), "d2h_sum": round(sum(x["d2h"] for x in rows), 1)}
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
