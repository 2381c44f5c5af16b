"""GPU parity tests proper: the CUDA path through the C ABI (libmfc5.so) against the CPU oracle and
against the reference's golden files.

* strict build (-fmad=false): per-stage RHS and 50-step state must agree with the oracle to a few ulp
  (same operation order, no FMA contraction) -- any larger difference is a logic bug;
* product build (FMA contraction + division-light weights / Riemann solve, ``fast_math = 1``, the default):
  50-step state must pass the reference's golden tolerance (1e-12 abs-or-rel) and agree with the
  oracle within 1e-10 max relative error (north-star bar); ``fast_math = 0`` (reference operation
  order, FMA only) is held to the same bar.
"""
import numpy as np
import pytest

from golden_util import CASES, ORDER7, VISCOUS, compare, golden, golden_field, setup_case
from mfc_b200 import driver

pytestmark = pytest.mark.gpu

SINGLE = sorted(u for u in CASES if CASES[u]["ppn"] == 1 and str(CASES[u]["params"].get("cfl_adap_dt", "F")) != "T")  # CFL-driven: test_gpu_driver.py
# one representative per feature for the (slower) oracle cross-check
ORACLE_SET = ["2AB32975", "677DAA82", "9090FC4B", "16A922E4", "C4907722", "1CCA82F5", "5EC236F2", "63CD6ADF", "70DAE9E8",
              "D6415F48", "AE3ECF01", "F0F175B2", "4129A23A", "EB6687AB", "10DBA76B", "4D7926CD", "55533234",
              "04E1CB26", "A0B82851", "5010B814", "B49877F3", "EFDCFF72", "758D0268", "D7107B89", "B5F0C655", "24FD489A"]


def _gpu(uuid, strict, fast=1):
    from mfc_b200.lib import Mfc5
    case, P, cbs, cons, dbuf = setup_case(uuid, fast_math=fast)
    g = Mfc5(P, cbs, dirichlet=dbuf, strict=strict)
    g.set_cons(cons)
    return case, P, cbs, cons, dbuf, g


def _oracle(P, cbs, cons, dbuf, omp=True):
    from oracle_lib import Oracle
    o = Oracle(P, cbs, omp=omp)
    for (d, side), buf in dbuf.items():
        o.set_dirichlet(d, side, buf)
    o.set_cons(cons)
    return o


def _maxrel(a, b):
    scale = np.maximum(np.abs(b), 1e-30)
    return float(np.max(np.abs(a - b) / scale))


@pytest.mark.parametrize("uuid", ORACLE_SET)
def test_rhs_strict_matches_oracle(uuid):
    """s_compute_rhs seam: primitive variables and RHS of the initial state, strict build vs oracle."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=True)
    o = _oracle(P, cbs, cons, dbuf)
    g.compute_rhs()
    o.compute_rhs()
    np.testing.assert_array_equal(g.get_prim(), o.get_prim())
    r_g, r_o = g.get_rhs(), o.get_rhs()
    scale = np.max(np.abs(r_o)) + 1e-300
    assert np.max(np.abs(r_g - r_o)) <= 1e-13 * scale, np.max(np.abs(r_g - r_o)) / scale


# the widening of SURVEY.md 8 f-4 (reference-order kernels): one or two goldens per feature
F4_SET = ["07C1FECC", "C462AA3E", "F9E0BE95", "7C8F1BA9", "A26B7E00", "4AADC848", "F1D01264", "01DC0251", "7BCDC8E7", "10041BB7",
          "516A2067", "A089CE54", "DC97B66D", "0ACB1F16", "A3EC9BD7", "EDADA16E"]


@pytest.mark.parametrize("uuid", F4_SET)
def test_f4_features_strict_vs_oracle(uuid):
    """low_Mach, alt_soundspeed, Lax-Friedrichs, MUSCL, TENO5, body forces, THINC, mp_weno: strict build == oracle, bit for bit, after 10
    steps.  THINC (516A2067, A089CE54) needs exp(): the oracle and the strict build evaluate the same fdlibm-style sequence
    (orc_exp / thinc_exp) -- with libm's exp on one side and CUDA's on the other, last-bit differences flip THINC's thresholds and
    show up at full size in the noise-level transverse momenta of these symmetric cases."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=True)
    o = _oracle(P, cbs, cons, dbuf)
    dt = float(case["dt"])
    for _ in range(10):
        g.step(dt)
        o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    assert not np.array_equal(b, cons)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("uuid", ORDER7)
def test_order7_strict_vs_oracle(uuid):
    """weno_order = 7 (WENO7-JS / -M / -Z, TENO7; m_weno.fpp:362-851 tables, :1248-1490 reconstruction) in the CUDA kernels: the strict
    build against the oracle after 10 steps -- bit for bit (tables from the same construction, same operation order), except
    WENO-Z, whose (tau / beta)^wenoz_q goes through pow(): libm on one side, CUDA's on the other (1e-13)."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=True)
    o = _oracle(P, cbs, cons, dbuf)
    dt = float(case["dt"])
    for _ in range(10):
        g.step(dt)
        o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    assert not np.array_equal(b, cons)
    if P.wenoz:
        assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b))
    else:
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("uuid", [u for u in VISCOUS if u in SINGLE])
def test_viscous_strict_vs_oracle(uuid):
    """viscous = T (finite-difference velocity gradients at the cell boundaries, m_viscous.fpp:515-830; viscous source flux
    m_riemann_state.fpp:805-1031; buff_size = 2 weno_polyn + 2): RHS of the initial state and the state after 10 steps, strict
    build vs oracle, bit for bit (1-D, 2-D, 3-D; one and two fluids)."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=True)
    assert g.b == 2 * ((P.weno_order - 1) // 2) + 2
    o = _oracle(P, cbs, cons, dbuf)
    g.compute_rhs()
    o.compute_rhs()
    np.testing.assert_array_equal(g.get_rhs(), o.get_rhs())
    dt = float(case["dt"])
    for _ in range(10):
        g.step(dt)
        o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    assert not np.array_equal(b, cons)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("uuid", ORACLE_SET)
def test_50_steps_vs_oracle(uuid, fast):
    """s_tvd_rk seam: after the case's 50 steps, product build within 1e-10 of the oracle (max rel)."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=False, fast=fast)
    o = _oracle(P, cbs, cons, dbuf)
    driver.run(g, case)
    driver.run(o, case)
    a, b = g.get_cons(), o.get_cons()
    tol = 1e-10 if CASES[uuid]["tol"] <= 1e-10 else CASES[uuid]["tol"]
    for v in range(a.shape[0]):
        ok, mabs, worst = compare(a[v], b[v], tol)
        assert ok, f"{uuid} var {v + 1}: worst min(abs, rel) err {worst:.3e}"


@pytest.mark.parametrize("uuid", SINGLE)
def test_golden(uuid):
    """The reference's own regression files, at the reference's own tolerance, through the C ABI."""
    case, P, cbs, cons, dbuf, g = _gpu(uuid, strict=False)
    driver.run(g, case)
    q = g.get_cons()
    gold = golden(uuid)
    shape = q.shape[1:]
    nstep = int(case["t_step_stop"])
    tol = CASES[uuid]["tol"]
    for v in range(q.shape[0]):
        ok, a, w = compare(q[v], golden_field(gold, "cons", v + 1, 0, nstep, shape), tol)
        assert ok, f"{uuid} ({CASES[uuid]['trace']}) var {v + 1}: worst min(abs,rel) err {w:.3e} > {tol}"
    assert g.kernel_launches > 0


def test_icfl_reduction_matches_oracle():
    case, P, cbs, cons, dbuf, g = _gpu("A0B82851", strict=True)
    o = _oracle(P, cbs, cons, dbuf)
    o.compute_rhs()
    dt = float(case["dt"])
    assert abs(g.icfl_max(dt) - o.icfl_max(dt)) <= 1e-14 * o.icfl_max(dt)
    assert g.reduce(2) == 0.0


# ---- the BASELINE.json configurations themselves (case-optimised kernels), reduced grids ----------------
def _baseline_cases():
    from mfc_b200 import cases
    return {
        "C5_perf_hllc_weno5m": cases.perf_test_3d(m=39, steps=10),
        "C5_perf_hll_weno5m": cases.perf_test_3d(m=39, riemann="hll", steps=10),
        "C4_droplet_weno3m": dict(cases.shockdroplet_3d(ny=30, steps=10)),
        "C4_droplet_weno5z": dict(cases.shockdroplet_3d(ny=30, weno_order=5, mapped_weno="F", wenoz="T", steps=10)),
        "C3_bench_weno3m": cases.bench_5eq_weno3_hllc(s=30, steps=10),
        "C2_shockbubble": dict(cases.shockbubble_2d(ny=40), t_step_stop=20),
        "C1_sod": dict(cases.sod_1d(), t_step_stop=100),
    }


@pytest.mark.parametrize("name", ["C5_perf_hllc_weno5m", "C5_perf_hll_weno5m", "C4_droplet_weno3m", "C4_droplet_weno5z",
                                  "C3_bench_weno3m", "C2_shockbubble", "C1_sod"])
@pytest.mark.parametrize("mode", ["strict", "fast", "ref_order"])
def test_baseline_config_vs_oracle(name, mode):
    """BASELINE configs (reduced grids, first steps) through the C ABI against the oracle:
    strict build bit-exact; product build (fast and reference-order arithmetic) within 1e-10 max relative
    error of max|q| per variable (north-star bar)."""
    from mfc_b200 import pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    from oracle_lib import Oracle
    case = _baseline_cases()[name]
    P = params_from_case(case, fast_math=1 if mode == "fast" else 0)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs, strict=(mode == "strict"))
    o = Oracle(P, cbs, omp=True)
    g.set_cons(cons)
    o.set_cons(cons)
    driver.run(g, case)
    driver.run(o, case)
    a, b = g.get_cons(), o.get_cons()
    assert not np.isnan(a).any()
    if mode == "strict":
        np.testing.assert_array_equal(a, b)
    else:
        for v in range(a.shape[0]):
            scale = np.max(np.abs(b[v])) + 1e-300
            err = np.max(np.abs(a[v] - b[v])) / scale
            assert err <= 1e-10, f"{name} var {v + 1}: {err:.3e}"
    g.close()


@pytest.mark.parametrize("name", ["C5_perf_hllc_weno5m", "C4_droplet_weno5z", "C3_bench_weno3m", "C2_shockbubble"])
def test_100_rk3_steps_within_1e10(name):
    """The north star's acceptance bar: conserved variables within 1e-10 max relative error after 100 RK3 steps in FP64,
    product build (division-light arithmetic, CUDA-graph replay) against the oracle restatement of the reference
    (the reference's own CPU build cannot be made in this image)."""
    from mfc_b200 import pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    from oracle_lib import Oracle
    case = dict(_baseline_cases()[name], t_step_stop=100, t_step_save=100)
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs)
    o = Oracle(P, cbs, omp=True)
    g.set_cons(cons)
    o.set_cons(cons)
    dt = float(case["dt"])
    g.steps(dt, 100)
    g.sync()
    for _ in range(100):
        o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    assert not np.isnan(a).any() and not np.array_equal(b, cons)
    worst = 0.0
    for v in range(a.shape[0]):
        scale = np.max(np.abs(b[v])) + 1e-300
        worst = max(worst, np.max(np.abs(a[v] - b[v])) / scale)
    # measured 2.4e-15 on C5 (tools/parity_growth.py); before the RK update divided exactly by 3 (div_by_const) the momenta
    # of this stiffened-gas case drifted to 6e-10 -- DESIGN.md section 4
    assert worst <= 1e-10, f"{name}: {worst:.3e}"
    g.close()


def _acceptance_cases():
    from mfc_b200 import cases
    return {
        # examples/3D_shockdroplet/case.py at its own size (201 x 101 x 101 cells, stretched grid), as shipped and with the north
        # star's WENO5-Z override
        "C4_full_weno3m": cases.shockdroplet_3d(ny=100, steps=100),
        "C4_full_weno5z": cases.shockdroplet_3d(ny=100, weno_order=5, mapped_weno="F", wenoz="T", steps=100),
        # examples/3D_performance_test at the bench size (201^3): ~5 minutes of oracle time on 16 cores, so only on request
        "C5_full_weno5m": cases.perf_test_3d(m=200, steps=100),
        # examples/2D_shockbubble (300 x 100 cells) and examples/1D_sodshocktube (400 cells) as shipped
        "C2_full": cases.shockbubble_2d(ny=100),
        "C1_full": cases.sod_1d(),
    }


@pytest.mark.parametrize("name", ["C4_full_weno3m", "C4_full_weno5z", "C5_full_weno5m", "C2_full", "C1_full"])
def test_full_size_100_rk3_steps_within_1e10(name):
    """The north star's acceptance bar AT THE STATED SIZES: 100 RK3 steps of the case file's own grid, product build (case-optimised
    kernels, CUDA-graph replay) against the oracle, max relative error of every conserved variable <= 1e-10.  The result line
    is appended to gpurun_out/acceptance.log (kept under profiles/).  C5 at 201^3 costs ~5 minutes of oracle time and runs when
    MFC5_ACCEPTANCE=1 (tools/gpu_accept.sh)."""
    import os
    import time
    from mfc_b200 import pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    from oracle_lib import Oracle
    if name.startswith("C5") and not os.environ.get("MFC5_ACCEPTANCE"):
        pytest.skip("set MFC5_ACCEPTANCE=1 (about 5 minutes of oracle time on 16 cores)")
    case = dict(_acceptance_cases()[name], t_step_stop=100, t_step_save=100)
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs)
    o = Oracle(P, cbs, omp=True)
    g.set_cons(cons)
    o.set_cons(cons)
    dt = float(case["dt"])
    t0 = time.time()
    g.steps(dt, 100)
    g.sync()
    t1 = time.time()
    for _ in range(100):
        o.step(dt)
    t2 = time.time()
    a, b = g.get_cons(), o.get_cons()
    g.close()
    assert not np.isnan(a).any() and not np.array_equal(b, cons)
    errs = [float(np.max(np.abs(a[v] - b[v])) / (np.max(np.abs(b[v])) + 1e-300)) for v in range(a.shape[0])]
    line = (f"{name}: cells {cons.shape[3]}x{cons.shape[2]}x{cons.shape[1]}, 100 RK3 steps, max rel err per variable "
            + " ".join(f"{e:.2e}" for e in errs) + f"; GPU {t1 - t0:.2f} s, oracle ({os.cpu_count()} threads) {t2 - t1:.1f} s")
    print(line)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/acceptance.log", "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
    assert max(errs) <= 1e-10, line


# ---- BASELINE.json's full size (examples/3D_performance_test, 201^3 cells): one step against the oracle, and the
# ---- size-independent properties of the scheme ----------------------------------------------------------------------
def _full_size_c5(fast, strict, **over):
    from mfc_b200 import cases, pre_process
    from mfc_b200.lib import Mfc5
    from mfc_b200.params import params_from_case
    case = dict(cases.perf_test_3d(m=200, steps=1), **over)
    P = params_from_case(case, fast_math=fast)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    g = Mfc5(P, cbs, strict=strict)
    return case, P, cbs, cons, g


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_full_size_c5_one_step_vs_oracle(mode):
    """The bench workload itself (201^3 cells, stretched grid, WENO5-M + HLLC + mpp_lim, BC -2/-6): one RK3 step through the C
    ABI against the oracle on the same input.  Strict build: bit for bit; product build (case-optimised kernels, CUDA
    graph): the reference's golden rule at 1e-12 (abs or rel), and 1e-13 of max|q| per variable."""
    from oracle_lib import Oracle
    case, P, cbs, cons, g = _full_size_c5(fast=0 if mode == "strict" else 1, strict=(mode == "strict"))
    o = Oracle(P, cbs, omp=True)
    g.set_cons(cons)
    o.set_cons(cons)
    dt = float(case["dt"])
    g.step(dt)
    o.step(dt)
    a, b = g.get_cons(), o.get_cons()
    g.close()
    assert not np.array_equal(b, cons)
    if mode == "strict":
        np.testing.assert_array_equal(a, b)
    else:
        for v in range(a.shape[0]):
            ok, mabs, worst = compare(a[v], b[v], 1e-12)
            assert ok, f"var {v + 1}: worst min(abs, rel) err {worst:.3e}"
            assert np.max(np.abs(a[v] - b[v])) <= 1e-13 * np.max(np.abs(b[v])) + 1e-300


@pytest.mark.parametrize("ends", ["extrapolation", "case"])
def test_full_size_uniform_state_is_preserved(ends):
    """Free-stream preservation at full size: a uniform two-fluid state at rest on the stretched 201^3 grid has equal fluxes
    on every face, so three RK3 steps leave it unchanged (product build; its WENO adds weighted increments to the cell
    value, which is exact for constant data).  With ghost-cell boundaries (-2 reflective, -3 extrapolation) the momenta stay
    exactly zero and the rest moves by at most the rounding of the RK combination (one ulp per step); with
    the case's own characteristic boundaries (-6) the boundary flux is an interpolation of cell-centre fluxes, which may
    round the pressure by an ulp: momenta stay below 1e-12, everything else is unchanged to 1e-14 relative."""
    over = {f"bc_{d}%end": -3 for d in "xyz"} if ends == "extrapolation" else {}
    case, P, cbs, cons, g = _full_size_c5(fast=1, strict=False, **over)
    q = np.empty_like(cons)
    for v in range(cons.shape[0]):
        q[v] = cons[v, -1, -1, -1]           # the far-field corner state everywhere
    nf = P.num_fluids
    q[nf:nf + 3] = 0.0                       # at rest
    g.set_cons(q)
    g.steps(float(case["dt"]), 3)
    g.sync()
    a = g.get_cons()
    g.close()
    if ends == "extrapolation":
        # no spurious motion at all: the momenta stay exactly zero; the other variables only see the rounding of the RK
        # combination itself ((2 q + q) / 3 is one ulp off q for q = 0.1 -- in the reference's arithmetic as well)
        assert not a[nf:nf + 3].any()
        assert np.all(np.abs(a - q) <= 3 * np.spacing(np.abs(q)))
    else:
        assert np.max(np.abs(a[nf:nf + 3])) <= 1e-12
        for v in list(range(nf)) + list(range(nf + 3, a.shape[0])):
            assert np.max(np.abs(a[v] - q[v])) <= 1e-14 * np.max(np.abs(q[v]))


def test_full_size_periodic_conservation():
    """Conservation at full size: with periodic boundaries on a uniform 201^3 grid the flux differences telescope, so the sums
    of the partial densities, momenta and energy over the domain change only by rounding (the volume fractions obey a
    non-conservative equation and are not checked).  Smooth two-fluid field, 3 RK3 steps, product build."""
    over = {"stretch_x": "F", "stretch_y": "F", "stretch_z": "F"}
    for d in "xyz":
        over[f"bc_{d}%beg"] = -1
        over[f"bc_{d}%end"] = -1
    case, P, cbs, cons, g = _full_size_c5(fast=1, strict=False, **over)
    nz, ny, nx = cons.shape[1:]
    z, y, x = np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij")
    two_pi = 2.0 * np.pi
    a1 = 0.5 + 0.3 * np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(two_pi * z)
    rho1, rho2 = 1000.0, 1.0
    u = [5.0 * np.sin(two_pi * (x + y)), 3.0 * np.cos(two_pi * (y - z)), 2.0 * np.sin(two_pi * (z + x))]
    pres = 1.0e5 * (1.0 + 0.1 * np.cos(two_pi * (x + y + z)))
    nf = P.num_fluids
    q = np.empty_like(cons)
    q[0], q[1] = a1 * rho1, (1.0 - a1) * rho2
    rho = q[0] + q[1]
    for d in range(3):
        q[nf + d] = rho * u[d]
    gam = a1 * P.fluid_gamma[0] + (1.0 - a1) * P.fluid_gamma[1]
    pinf = a1 * P.fluid_pi_inf[0] + (1.0 - a1) * P.fluid_pi_inf[1]
    q[nf + 3] = gam * pres + pinf + 0.5 * rho * (u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
    q[nf + 4], q[nf + 5] = a1, 1.0 - a1
    g.set_cons(q)
    g.steps(float(case["dt"]), 3)
    g.sync()
    a = g.get_cons()
    g.close()
    assert not np.isnan(a).any() and not np.array_equal(a, q)
    for v in range(nf + 4):
        s0, s1 = np.sum(q[v], dtype=np.longdouble), np.sum(a[v], dtype=np.longdouble)
        scale = np.sum(np.abs(q[v]), dtype=np.longdouble) + 1e-300
        assert abs(s1 - s0) <= 1e-13 * scale, f"var {v + 1}: sum drifted by {float(abs(s1 - s0) / scale):.3e}"


@pytest.mark.parametrize("uuid", ["2AB32975", "4129A23A", "A0B82851"])
def test_opt_in_fused_cons2prim_is_bitwise(uuid, monkeypatch):
    """MFC5_FUSE_C2P=1 (the RK-fused last sweep also writes the primitives of the new state; second primitive buffer,
    graphs keyed by both buffer parities): same bits as the default path, eager and graph replay, strict build vs oracle."""
    case, P, cbs, cons, dbuf, g0 = _gpu(uuid, strict=True)
    monkeypatch.setenv("MFC5_FUSE_C2P", "1")
    _, _, _, _, _, g1 = _gpu(uuid, strict=True)
    _, _, _, _, _, g2 = _gpu(uuid, strict=True)
    o = _oracle(P, cbs, cons, dbuf)
    dt = float(case["dt"])
    for _ in range(6):
        g0.step(dt)
        g1.step(dt)
        o.step(dt)
    g2.steps(dt, 6)
    g2.sync()
    a = g0.get_cons()
    np.testing.assert_array_equal(a, o.get_cons())
    np.testing.assert_array_equal(g1.get_cons(), a)
    np.testing.assert_array_equal(g2.get_cons(), a)
    assert g1.icfl_max(dt) == g0.icfl_max(dt)
    g1.compute_rhs()
    g0.compute_rhs()
    np.testing.assert_array_equal(g1.get_rhs(), g0.get_rhs())


@pytest.mark.parametrize("uuid", ["2AB32975", "4129A23A", "B5F0C655"])
def test_graph_replay_is_bitwise_eager(uuid):
    """mfc5_tvd_rk_n (CUDA-graph replay of the captured step) == the same number of mfc5_tvd_rk calls."""
    case, P, cbs, cons, dbuf, g1 = _gpu(uuid, strict=False)
    _, _, _, _, _, g2 = _gpu(uuid, strict=False)
    dt = float(case["dt"])
    l1 = g1.kernel_launches  # the upload's repack kernel is already counted
    for _ in range(7):
        g1.step(dt)
    l0 = g2.kernel_launches
    g2.steps(dt, 7)
    g2.sync()
    assert g2.kernel_launches - l0 == g1.kernel_launches - l1
    np.testing.assert_array_equal(g1.get_cons(), g2.get_cons())
