"""world_size-2 run over torch.distributed (gloo, CPU): each process owns one block of the decomposition
(mfc_b200/decomp.py), halo faces travel by dist.send/recv in the same order the CUDA library issues its
ncclSend/ncclRecv (sends low face then high face, receives high then low, x then y then z), the arithmetic
is the CPU oracle.  Checked against the reference's own 2-rank goldens (per-rank files cons.<v>.<rank>)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import CASES, compare, golden

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange(o, blk, d):
    """Non-blocking pairs so the pattern is deadlock free for any neighbour graph, like a grouped NCCL call."""
    reqs, recv = [], {}
    for side in range(2):
        nb = blk.bc[d][side]
        if nb >= 0:
            buf = torch.from_numpy(o.halo_pack(d, side))
            reqs.append(dist.isend(buf, dst=nb, tag=d * 2 + side))
    for side in (1, 0):
        nb = blk.bc[d][side]
        if nb >= 0:
            recv[side] = torch.empty(o.L.orc_halo_count(o.h, d), dtype=torch.float64)
            # the neighbour sent this face as its (1 - side) face
            reqs.append(dist.irecv(recv[side], src=nb, tag=d * 2 + (1 - side)))
    for r in reqs:
        r.wait()
    for side, buf in recv.items():
        o.halo_unpack(d, side, buf.numpy())


def _worker(rank, world, port, uuid, q):
    import sys
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfc_b200 import decomp, driver, pre_process
    from mfc_b200.params import params_from_case
    from oracle_lib import Oracle
    case = CASES[uuid]["params"]
    blk = decomp.Block(case, rank, world)
    lc = blk.local_case(case)
    P = params_from_case(lc)
    cbs, ccs, gws = blk.grid(case, P.buff_size)
    _, cons = pre_process.initial_condition(lc, ccs)
    o = Oracle(P, cbs, ghost_widths=gws)
    o.set_cons(cons)
    for dt in driver.time_steps(case):
        for st in range(1, int(case.get("time_stepper", 3)) + 1):
            o.cons2prim()
            for d in range(blk.nd):
                for side in range(2):
                    if blk.bc[d][side] < 0:
                        o.bc_fill(d, side)
                _exchange(o, blk, d)
            o.sweeps()
            o.rk_update(st, dt)
    # global ICFL the way the library reduces it (max over ranks)
    icfl = torch.tensor([o.icfl_max(float(case["dt"]))], dtype=torch.float64)
    dist.all_reduce(icfl, op=dist.ReduceOp.MAX)
    q.put((rank, o.get_cons(), float(icfl.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("uuid", ["0FCCE9F1", "8C7AA13B", "CE232828"])
def test_two_ranks_over_gloo(uuid):
    world = CASES[uuid]["ppn"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, uuid, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, cons, icfl = q.get(timeout=600)
        out[r] = (cons, icfl)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = CASES[uuid]["params"]
    g, nstep, tol = golden(uuid), int(case["t_step_stop"]), CASES[uuid]["tol"]
    assert out[0][1] == out[1][1] and 0.0 < out[0][1] < 1.0
    for r in range(world):
        a = out[r][0]
        nz, ny, nx = a.shape[1:]
        for v in range(a.shape[0]):
            ref = g[f"cons.{v + 1}.{r:02d}.{nstep:06d}"].reshape(nx, ny, nz).transpose(2, 1, 0)
            ok, _, w = compare(a[v], ref, tol)
            assert ok, f"{uuid} rank {r} var {v + 1}: {w:.3e}"
