"""Multi-GPU parity (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): one process per GPU, block
decomposition as the reference, halo faces by ncclSend/ncclRecv inside libmfc5.so.  Checked against the
reference's 2-rank goldens and against the single-GPU run of the same case."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from golden_util import CASES, compare, golden

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, uid, case, q, strict):
    import sys
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from mfc_b200 import decomp, driver, pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    blk = decomp.Block(case, rank, world)
    lc = blk.local_case(case)
    P = params_from_case(lc, device=rank)
    cbs, ccs, gws = blk.grid(case, P.buff_size)
    _, cons = pre_process.initial_condition(lc, ccs, ds_min=pre_process.grid_minima(case))
    g = Mfc5(P, cbs, ghost_widths=gws, strict=strict)
    g.comm_init(uid, rank, world)
    g.set_cons(cons)
    driver.run(g, case)
    icfl = g.icfl_max(float(case["dt"]))
    q.put((rank, blk.start, g.get_cons(), icfl))
    g.close()


def _run(case, world, strict=False):
    from mfc_b200 import lib
    uid = lib.nccl_unique_id(strict)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, case, q, strict)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, start, cons, icfl = q.get(timeout=600)
        out[r] = (start, cons, icfl)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("uuid", ["0FCCE9F1", "8C7AA13B", "CE232828"])
def test_two_gpu_goldens(uuid):
    _need(2)
    case = CASES[uuid]["params"]
    out = _run(case, 2)
    g, nstep, tol = golden(uuid), int(case["t_step_stop"]), CASES[uuid]["tol"]
    assert out[0][2] == out[1][2]  # all-reduced ICFL
    for r in range(2):
        a = out[r][1]
        nz, ny, nx = a.shape[1:]
        for v in range(a.shape[0]):
            ref = g[f"cons.{v + 1}.{r:02d}.{nstep:06d}"].reshape(nx, ny, nz).transpose(2, 1, 0)
            ok, _, w = compare(a[v], ref, tol)
            assert ok, f"{uuid} rank {r} var {v + 1}: {w:.3e}"


@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposed_equals_single_gpu(world):
    """3D_performance_test (reduced), decomposed over `world` GPUs vs the single-GPU run.  Halo values are
    copies of the neighbour's primitives; the only difference is the ghost cell-boundary
    coordinates, which each rank rebuilds from the neighbour's cell widths as the reference does
    (m_mpi_common.fpp:1745-1796) -- an ulp-level change of the WENO tables on a stretched grid."""
    _need(world)
    from mfc_b200 import cases, driver, pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    case = dict(cases.perf_test_3d(m=59, steps=5), parallel_io="T")
    out = _run(case, world)
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs)
    g.set_cons(cons)
    driver.run(g, case)
    full = g.get_cons()
    for r in range(world):
        start, a, _ = out[r]
        nz, ny, nx = a.shape[1:]
        ref = full[:, start[2]:start[2] + nz, start[1]:start[1] + ny, start[0]:start[0] + nx]
        for v in range(a.shape[0]):
            err = np.abs(a[v] - ref[v]).max() / (np.abs(full[v]).max() + 1e-300)
            assert err <= 1e-12, f"rank {r} var {v + 1}: {err:.3e}"


@pytest.mark.parametrize("world", [2, 4, 8])
def test_viscous_decomposed_equals_single_gpu(world):
    """viscous = T on processor boundaries: the velocity gradients of a face read edge / corner ghost cells, so the ghost region
    is filled and exchanged direction by direction over the ghost layers of the earlier directions (x, then y over the x ghosts,
    then z over both; m_boundary_common.fpp:99-187, m_mpi_common.fpp:620-1065).  The reference's 3-D two-fluid viscous test case
    (1C0780C8: shear and bulk Reynolds numbers, uniform grid) on 50^3 cells, decomposed vs the single-GPU run, strict build: the
    single-GPU strict build is itself bit-identical with the oracle (tests/test_gpu_parity.py::test_viscous_strict_vs_oracle)."""
    _need(world)
    from mfc_b200 import driver, pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    case = dict(CASES["1C0780C8"]["params"], m=49, n=49, p=49, t_step_stop=5, t_step_save=5, parallel_io="T")
    out = _run(case, world, strict=True)
    P = params_from_case(case)
    assert P.viscous == 1
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs, strict=True)
    g.set_cons(cons)
    driver.run(g, case)
    full = g.get_cons()
    g.close()
    assert not np.array_equal(full, cons)
    for r in range(world):
        start, a, _ = out[r]
        nz, ny, nx = a.shape[1:]
        ref = full[:, start[2]:start[2] + nz, start[1]:start[1] + ny, start[0]:start[0] + nx]
        for v in range(a.shape[0]):
            err = np.abs(a[v] - ref[v]).max() / (np.abs(full[v]).max() + 1e-300)
            assert err <= 1e-13, f"rank {r} var {v + 1}: {err:.3e}"
