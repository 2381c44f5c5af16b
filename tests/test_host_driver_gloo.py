"""The host driver (mfc_b200/main.py: pre_process + p_main loop + on-disk formats) on 2 ranks over gloo, CPU only:
each process is one rank of ``torchrun -m mfc_b200 run`` with the CPU oracle + gloo halo exchange standing in for the
CUDA library + NCCL (test infrastructure).  Checked against the reference's 2-rank goldens through the per-rank D/ files,
and parallel_io = T (one restart_data file written by both ranks) against the serial-I/O run."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

from golden_util import CASES, compare, golden
from test_decomp_gloo import _exchange, _free_port

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, case_dir, q):
    import sys
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from mfc_b200 import decomp, main
    from oracle_lib import Oracle

    class GlooOracle(Oracle):
        """step() = the stage sequence of the CUDA library with dist.send/recv in place of ncclSend/ncclRecv."""

        def step(self, dt):
            for st in range(1, self.P.time_stepper + 1):
                self.cons2prim()
                for d in range(self.blk.nd):
                    for side in range(2):
                        if self.blk.bc[d][side] < 0:
                            self.bc_fill(d, side)
                    _exchange(self, self.blk, d)
                self.sweeps()
                self.rk_update(st, dt)

        def icfl_max(self, dt):  # primitives of the current state, maximum over ranks (the library: ncclAllReduce)
            import torch
            self.cons2prim()
            t = torch.tensor([Oracle.icfl_max(self, dt)], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

    w = main.World()  # initialises the gloo group from the environment, like under torchrun
    blk = decomp.Block(case, rank, world)

    def make(P, cbs, gws, dirichlet):
        o = GlooOracle(P, cbs, ghost_widths=gws)
        o.blk = blk
        return o
    main.run_pre_process(case, case_dir, w, log=lambda s: None)
    out = main.run_simulation(case, case_dir, w, make_state=make, log=lambda s: None)
    q.put((rank, out["cons"], out["grind_time"]))
    dist.barrier()
    dist.destroy_process_group()


def _run(case, case_dir, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, case_dir, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            r, cons, grind = q.get(timeout=240)
            out[r] = cons
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:  # a rank that failed leaves its peer blocked in a receive
            if p.is_alive():
                p.terminate()
    return out


@pytest.mark.parametrize("uuid", ["0FCCE9F1", "8C7AA13B"])
def test_two_rank_driver_matches_golden(uuid, tmp_path):
    from mfc_b200 import data_io
    case = dict(CASES[uuid]["params"])
    _run(case, str(tmp_path))
    packed = data_io.pack_D(str(tmp_path))
    g = golden(uuid)
    assert any(".01." in k for k in g)
    for key, want in g.items():
        name = f"D/{key}.dat"
        assert name in packed, name
        ok, amax, worst = compare(packed[name], want, float(CASES[uuid]["tol"]))
        assert ok, (uuid, key, amax, worst)
    assert os.path.isdir(tmp_path / "p_all" / "p1" / "50")
    assert open(tmp_path / "time_data.dat").read().split("\n")[1].split()[0] == "2"


def test_two_rank_parallel_io(tmp_path):
    """parallel_io = T: both ranks write their block of restart_data/lustre_<t>.dat.  Against the serial-I/O run only to the
    reference tolerance: with serial I/O every rank builds its grid from its own domain bounds (m_mpi_common.fpp:1681-1740),
    with parallel I/O it slices the global grid (:1800), and the cell boundaries differ in the last bits."""
    from mfc_b200 import data_io
    case = dict(CASES["8C7AA13B"]["params"])
    a = _run(dict(case, parallel_io="T"), str(tmp_path / "par"))
    b = _run(dict(case, parallel_io="F"), str(tmp_path / "ser"))
    for r in (0, 1):
        ok, amax, worst = compare(a[r], b[r], 1e-12)
        assert ok, (r, amax, worst)
    m, n, p = int(case["m"]), int(case["n"]), int(case.get("p", 0))
    nsys = a[0].shape[0]
    f = tmp_path / "par" / "restart_data" / "lustre_50.dat"
    assert os.path.getsize(f) == nsys * (m + 1) * (n + 1) * (p + 1) * 8
    glb = np.fromfile(f).reshape(nsys, p + 1, n + 1, m + 1)
    from mfc_b200 import decomp
    for r in (0, 1):
        blk = decomp.Block(case, r, 2)
        assert np.array_equal(blk.slab(glb), a[r])
