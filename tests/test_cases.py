"""The BASELINE configurations (mfc_b200/cases.py) against the parameter dictionaries the reference's own
test generator expands for its example-based goldens (tests/golden/cases.json, made by make_fixtures.py)."""
import numpy as np
import pytest

from golden_util import CASES
from mfc_b200 import cases, pre_process
from mfc_b200.params import _NAMED, params_from_case


def _norm(k, v):
    return _NAMED[k][v] if isinstance(v, str) and k in _NAMED else v


@pytest.mark.parametrize("builder,uuid,skip", [
    (cases.sod_1d, "BCA926AC", ("t_step_stop", "t_step_save", "parallel_io")),
    (cases.shockbubble_2d, "24FD489A", ("t_step_stop", "t_step_save", "m", "n", "parallel_io")),
    (cases.shockdroplet_3d, "B5F0C655", ("t_step_stop", "t_step_save", "m", "n", "p", "parallel_io")),
])
def test_builder_matches_reference_expansion(builder, uuid, skip):
    # (the test generator turns parallel_io off so that D/ is written; the stock files have it on)
    mine, ref = builder(), CASES[uuid]["params"]
    for k, v in mine.items():
        if k in skip or (k not in ref and v == "F"):
            continue
        assert k in ref, k
        assert _norm(k, v) == _norm(k, ref[k]), (k, v, ref[k])


def test_benchmark_size_rule():
    # benchmarks/5eq_rk3_weno3_hllc/case.py:24-28: --gbpp 16 on one rank -> s = floor((8e6/2)^(1/3)) = 158
    c = cases.bench_5eq_weno3_hllc(gbpp=16)
    assert (c["m"], c["n"], c["p"]) == (316, 158, 158)
    c = cases.bench_5eq_weno3_hllc(gbpp=180)
    assert (c["m"], c["n"], c["p"]) == (710, 355, 355)


def test_perf_test_initial_condition_is_smoothed_and_bounded():
    case = cases.perf_test_3d(m=31)
    P = params_from_case(case)
    assert (P.weno_order, P.mapped_weno, P.riemann_solver, P.mpp_lim, P.sys_size) == (5, 1, 2, 1, 8)
    cbs, ccs = pre_process.generate_grid(case)
    prim, cons = pre_process.initial_condition(case, ccs)
    a1 = prim[6]
    assert 0.1 - 1e-12 <= a1.min() and a1.max() <= 0.9 + 1e-12
    assert len(np.unique(np.round(a1, 6))) > 10  # tanh profile, not a step
    np.testing.assert_allclose(prim[6] + prim[7], 1.0, atol=1e-14)
    assert np.all(np.diff(cbs[0]) > 0) and abs(np.diff(cbs[0]).max() / np.diff(cbs[0]).min()) > 1.5  # stretched


def test_analytic_patch_expression_and_hardcoded_profile():
    """Analytic patch values (toolchain/mfc/case.py:152-255) evaluated over the patch's cells; hcid 180 is the same
    Shu-Osher profile with the reference's single-precision 0.2 (src/common/include/1dHardcodedIC.fpp:21-26)."""
    base = dict(CASES["B50DE910"]["params"])
    cbs, ccs = pre_process.generate_grid(base)
    prim_h, _ = pre_process.initial_condition(base, ccs)
    ana = dict(base)
    del ana["patch_icpp(2)%hcid"]
    ana["patch_icpp(2)%alpha_rho(1)"] = "1 + 0.2*sin(5*x)"
    prim_a, _ = pre_process.initial_condition(ana, ccs)
    x = ccs[0]
    inside = np.abs(x - float(base["patch_icpp(2)%x_centroid"])) <= 0.5 * float(base["patch_icpp(2)%length_x"])
    np.testing.assert_array_equal(prim_a[0][0, 0, inside], 1 + 0.2 * np.sin(5 * x[inside]))
    assert np.max(np.abs(prim_a[0] - prim_h[0])) < 4e-9 and np.max(np.abs(prim_a[0] - prim_h[0])) > 1e-10
    np.testing.assert_array_equal(prim_a[0][0, 0, ~inside], prim_h[0][0, 0, ~inside])
    with pytest.raises(ValueError):
        pre_process.initial_condition(dict(ana, **{"patch_icpp(2)%alpha_rho(1)": "__import__('os').getcwd()"}), ccs)
    with pytest.raises(ValueError):
        pre_process.initial_condition(dict(ana, **{"patch_icpp(2)%alpha_rho(1)": "1 + q"}), ccs)


def test_elliptic_smoothing_is_conservative_filter():
    case = dict(CASES["8C1E22E4"]["params"])
    cbs, ccs = pre_process.generate_grid(case)
    prim_s, _ = pre_process.initial_condition(case, ccs)
    prim_0, _ = pre_process.initial_condition(dict(case, elliptic_smoothing="F"), ccs)
    assert prim_s.shape == prim_0.shape and not np.array_equal(prim_s, prim_0)
    for v in range(prim_0.shape[0]):  # a convex filter cannot create new extrema
        assert prim_s[v].min() >= prim_0[v].min() - 1e-14 and prim_s[v].max() <= prim_0[v].max() + 1e-14


def test_rank_affinity_helper_is_safe_without_a_gpu():
    """mfc_b200/affinity.py (ranks next to their GPU before they allocate pinned memory): without NVML / a GPU it reports why it
    did nothing and leaves the process's CPU mask alone."""
    import os
    from mfc_b200 import affinity
    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    info = affinity.bind_to_gpu(0, 1)
    assert info["gpu"] == 0 and "bound" in info and info["numa_nodes"] >= 1
    if not info["bound"]:
        assert "note" in info
        if before is not None:
            assert os.sched_getaffinity(0) == before
    else:   # a GPU box: the mask is a non-empty subset of what the process was allowed
        assert before is None or os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)
