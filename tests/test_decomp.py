"""Block decomposition (mfc_b200/decomp.py) against the reference's rule and its 2-rank goldens."""
import numpy as np
import pytest

from golden_util import CASES, compare, golden
from mfc_b200 import decomp, driver, pre_process
from mfc_b200.params import UnsupportedCase, params_from_case
from oracle_lib import Oracle

TWO_RANK = ["0FCCE9F1", "8C7AA13B", "CE232828"]


def test_factorisation_rule():
    # 201x101x101 (3D_shockdroplet) with the reference's real-valued score, last minimum wins
    assert decomp.factorize(2, 200, 100, 100, 3) == (2, 1, 1)
    assert decomp.factorize(4, 200, 100, 100, 3) == (4, 1, 1)
    assert decomp.factorize(8, 200, 100, 100, 3) == (2, 2, 2)
    assert decomp.factorize(8, 401, 401, 401, 5) == (2, 2, 2)
    assert decomp.factorize(2, 29, 29, 49, 5) == (1, 1, 2)      # tests/CE232828 is split in z
    assert decomp.factorize(2, 49, 39, 0, 5) == (2, 1, 1)
    with pytest.raises(UnsupportedCase):
        decomp.factorize(8, 24, 24, 24, 5)                       # fewer than 25 cells per rank per direction


def test_blocks_tile_the_domain_and_link_neighbours():
    case = dict(CASES["6FC6A809"]["params"])  # 3-D periodic
    case.update(m=63, n=49, p=52)
    nr = 8
    blocks = [decomp.Block(case, r, nr, dims=(2, 2, 2)) for r in range(nr)]
    seen = np.zeros((53, 50, 64), dtype=int)
    for b in blocks:
        seen[b.start[2]:b.start[2] + b.s[2] + 1, b.start[1]:b.start[1] + b.s[1] + 1, b.start[0]:b.start[0] + b.s[0] + 1] += 1
    assert (seen == 1).all()
    for b in blocks:
        for d in range(3):
            for side in range(2):
                nb = b.bc[d][side]
                assert nb >= 0  # periodic everywhere -> every face is a processor boundary
                assert blocks[nb].bc[d][1 - side] == b.rank
    # remainder cells go to the low ranks (m_mpi_common.fpp:1482-1493)
    assert decomp.split(52, 2, 0) == (0, 26) and decomp.split(52, 2, 1) == (27, 25)


def run_ranks_in_process(case, nr, on_exchange=None):
    """All ranks of a decomposed run inside one process; halo faces handed over by the oracle's own
    pack/unpack (the CPU statement of s_mpi_sendrecv_variables_buffers)."""
    blocks = [decomp.Block(case, r, nr) for r in range(nr)]
    orcs = []
    for b in blocks:
        lc = b.local_case(case)
        P = params_from_case(lc)
        cbs, ccs, gws = b.grid(case, P.buff_size)
        _, cons = pre_process.initial_condition(lc, ccs)
        o = Oracle(P, cbs, ghost_widths=gws)
        o.set_cons(cons)
        orcs.append(o)
    nstage = int(case.get("time_stepper", 3))
    for dt in driver.time_steps(case):
        for st in range(1, nstage + 1):
            for o in orcs:
                o.cons2prim()
            for d in range(blocks[0].nd):
                for o, b in zip(orcs, blocks):
                    for side in range(2):
                        if b.bc[d][side] < 0:
                            o.bc_fill(d, side)
                bufs = {(b.rank, side): o.halo_pack(d, side) for o, b in zip(orcs, blocks) for side in range(2) if b.bc[d][side] >= 0}
                for o, b in zip(orcs, blocks):
                    for side in range(2):
                        if b.bc[d][side] >= 0:
                            o.halo_unpack(d, side, bufs[(b.bc[d][side], 1 - side)])
            for o in orcs:
                o.sweeps()
                o.rk_update(st, dt)
    return blocks, orcs


@pytest.mark.parametrize("uuid", TWO_RANK)
def test_two_rank_goldens(uuid):
    case = CASES[uuid]["params"]
    blocks, orcs = run_ranks_in_process(case, CASES[uuid]["ppn"])
    g, nstep, tol = golden(uuid), int(case["t_step_stop"]), CASES[uuid]["tol"]
    for r, o in enumerate(orcs):
        q = o.get_cons()
        nz, ny, nx = q.shape[1:]
        for v in range(q.shape[0]):
            ref = g[f"cons.{v + 1}.{r:02d}.{nstep:06d}"].reshape(nx, ny, nz).transpose(2, 1, 0)
            ok, _, w = compare(q[v], ref, tol)
            assert ok, f"{uuid} rank {r} var {v + 1}: {w:.3e}"


def test_decomposed_run_is_bitwise_identical_to_single_rank():
    """halo values are copies of the neighbour's primitives, so an N-rank run reproduces the
    1-rank run bit for bit (uniform grid, the global grid slice is used on every rank)."""
    case = dict(CASES["D6415F48"]["params"], parallel_io="T")  # 2-D WENO5-M
    case["t_step_stop"] = 5
    blocks, orcs = run_ranks_in_process(case, 2)
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    o = Oracle(P, cbs)
    o.set_cons(cons)
    driver.run(o, case)
    full = o.get_cons()
    for b, oo in zip(blocks, orcs):
        assert np.array_equal(oo.get_cons(), b.slab(full))
