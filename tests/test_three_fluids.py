"""num_fluids = 3 (the library accepts 1..3; the reference's suite holds no 3-fluid golden for the 5-equation model).

Empty-fluid invariance pins that path to the 2-fluid goldens: adding a third fluid with the EOS of fluid 2 and zero volume
fraction / partial density everywhere adds exact zeros to every mixture sum, so the shared variables must come out BIT FOR
BIT as in the 2-fluid run (same operation order), and alpha_3 / alpha_rho_3 must stay exactly zero.  Checked on the oracle
(CPU), on the strict CUDA build against the 3-fluid oracle, and on the product build against the reference's 2-fluid golden
files at their own tolerance.  N = 2 NF + ND + 1 is 10 in 3-D: the sweep kernels' N > 8 configuration (one marching block per
SM, single-buffered x window) is exercised only here.
"""
import numpy as np
import pytest

from golden_util import CASES, compare, golden, golden_field
from mfc_b200 import driver, pre_process
from mfc_b200.params import params_from_case

UUIDS = ["C4907722", "F0F175B2", "A6AC2E06", "A0B82851", "AACF1BC5", "851F7AE2"]  # 1-D / 2-D / 3-D, HLLC and HLL, mpp_lim, Roe


def with_empty_third_fluid(case):
    c = dict(case)
    c["num_fluids"] = 3
    for key in ("gamma", "pi_inf", "qv", "cv", "qvp"):
        if f"fluid_pp(2)%{key}" in c:
            c[f"fluid_pp(3)%{key}"] = c[f"fluid_pp(2)%{key}"]
    for i in range(1, int(c["num_patches"]) + 1):
        c[f"patch_icpp({i})%alpha_rho(3)"] = 0.0
        c[f"patch_icpp({i})%alpha(3)"] = 0.0
    return c


def _setup(case, **kw):
    P = params_from_case(case, **kw)
    cbs, ccs = pre_process.generate_grid(case)
    prim, cons = pre_process.initial_condition(case, ccs)
    return P, cbs, cons, pre_process.dirichlet_buffers(P, prim)


def _shared(nd):
    """index pairs (3-fluid, 2-fluid) of the variables both layouts hold"""
    three = [0, 1] + [3 + d for d in range(nd)] + [3 + nd] + [4 + nd, 5 + nd]
    two = [0, 1] + [2 + d for d in range(nd)] + [2 + nd] + [3 + nd, 4 + nd]
    return list(zip(three, two))


def _oracle(P, cbs, cons, dbuf, omp):
    from oracle_lib import Oracle
    o = Oracle(P, cbs, omp=omp)
    for (d, side), buf in dbuf.items():
        o.set_dirichlet(d, side, buf)
    o.set_cons(cons)
    return o


@pytest.mark.parametrize("uuid", UUIDS)
def test_oracle_empty_third_fluid_is_invisible(uuid):
    case2 = CASES[uuid]["params"]
    case3 = with_empty_third_fluid(case2)
    out = []
    for case in (case2, case3):
        P, cbs, cons, dbuf = _setup(case)
        o = _oracle(P, cbs, cons, dbuf, omp=P.num_dims == 3)
        driver.run(o, case)
        out.append((P, o.get_cons()))
    (P2, q2), (P3, q3) = out
    assert P3.sys_size == P2.sys_size + 2
    nd = P2.num_dims
    for i3, i2 in _shared(nd):
        np.testing.assert_array_equal(q3[i3], q2[i2])
    assert not q3[2].any() and not q3[5 + nd + 1].any()


@pytest.mark.gpu
@pytest.mark.parametrize("uuid", UUIDS)
def test_gpu_three_fluids_strict_vs_oracle(uuid):
    """strict CUDA build == oracle, bit for bit, on the 3-fluid case (10 steps)"""
    from mfc_b200.lib import Mfc5
    case3 = with_empty_third_fluid(CASES[uuid]["params"])
    P, cbs, cons, dbuf = _setup(case3, fast_math=0)
    g = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    g.set_cons(cons)
    o = _oracle(P, cbs, cons, dbuf, omp=True)
    dt = float(case3["dt"])
    for _ in range(10):
        g.step(dt)
        o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    g.close()
    assert not np.array_equal(b, cons)
    np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("uuid", UUIDS)
def test_gpu_three_fluids_vs_two_fluid_golden(uuid, fast):
    """product build on the 3-fluid case: the shared variables pass the reference's 2-fluid golden at its tolerance"""
    from mfc_b200.lib import Mfc5
    case3 = with_empty_third_fluid(CASES[uuid]["params"])
    P, cbs, cons, dbuf = _setup(case3, fast_math=fast)
    g = Mfc5(P, cbs, dirichlet=dbuf)
    g.set_cons(cons)
    driver.run(g, case3)
    q = g.get_cons()
    g.close()
    gold = golden(uuid)
    nd = P.num_dims
    nstep = int(case3["t_step_stop"])
    for i3, i2 in _shared(nd):
        ok, a, w = compare(q[i3], golden_field(gold, "cons", i2 + 1, 0, nstep, q.shape[1:]), CASES[uuid]["tol"])
        assert ok, f"{uuid} var {i2 + 1}: worst min(abs, rel) err {w:.3e}"
    assert np.max(np.abs(q[2])) <= 1e-300 and np.max(np.abs(q[5 + nd + 1])) <= 1e-300


# ---- the same invariance one level down: an empty SECOND fluid on the single-fluid boundary-condition goldens ----------------
# The reference's goldens for the characteristic (-5 .. -12, GRCBC), slip / no-slip (-15 / -16) and Dirichlet (-17) boundaries are
# single-fluid, while every BASELINE configuration that uses them (C2 - C5: -6) has two fluids.  The 2-fluid CUDA path is held
# bit for bit to the 2-fluid oracle on those configurations (tests/test_gpu_parity.py::test_baseline_config_vs_oracle); this test
# closes the chain to the goldens: the 2-fluid oracle with an empty second fluid == the golden-pinned 1-fluid oracle, bit for bit.
BC_UUIDS = ["727F72ED", "A60691E7", "3FC6FC4A", "2FD933A2", "ED728400", "9D92767D", "3AE495F4", "5EC236F2", "48CCE072", "AED93D34",
            "C5B79059",                                                                # 1-D: -10 -11 -12 -15 -16 -17 -5 .. -9
            "B96AC58F", "4D7926CD", "98081DA3", "10DBA76B", "A6E65782", "4129A23A", "E84967E7", "5F877BC9", "16C03D8E",  # 2-D
            "9DCBE54F", "5010B814", "730DFD6D"]                                        # 3-D: -17, -6, -7
# (GRCBC inflow is NOT invariant, in the reference itself: every fluid's wave amplitude carries the pressure-relaxation term
#  - c Ma (p - p_in) / Del whatever its partial density, m_cbc.fpp:739-742, so an empty fluid is filled at the inflow face.)


def with_empty_second_fluid(case):
    c = dict(case)
    c["num_fluids"] = 2
    for key in ("gamma", "pi_inf", "qv", "cv", "qvp"):
        if f"fluid_pp(1)%{key}" in c:
            c[f"fluid_pp(2)%{key}"] = c[f"fluid_pp(1)%{key}"]
    for i in range(1, int(c["num_patches"]) + 1):
        c[f"patch_icpp({i})%alpha_rho(2)"] = 0.0
        c[f"patch_icpp({i})%alpha(2)"] = 0.0
    return c


@pytest.mark.parametrize("uuid", BC_UUIDS)
def test_oracle_empty_second_fluid_on_boundary_goldens(uuid):
    case1 = CASES[uuid]["params"]
    case2 = with_empty_second_fluid(case1)
    out = []
    for case in (case1, case2):
        P, cbs, cons, dbuf = _setup(case)
        o = _oracle(P, cbs, cons, dbuf, omp=P.num_dims == 3)
        driver.run(o, case)
        out.append((P, o.get_cons()))
    (P1, q1), (P2, q2) = out
    nd = P1.num_dims
    # 1-fluid layout: [arho, mom(nd), E, alpha]; 2-fluid: [arho1, arho2, mom(nd), E, alpha1, alpha2]
    pairs = [(0, 0)] + [(2 + d, 1 + d) for d in range(nd)] + [(2 + nd, 1 + nd), (3 + nd, 2 + nd)]
    for i2, i1 in pairs:
        np.testing.assert_array_equal(q2[i2], q1[i1])
    assert not q2[1].any() and not q2[4 + nd].any()
