"""The reference's known-answer convergence suite (toolchain/mfc/test/cases.py:25-175, toolchain/mfc/test/convergence.py),
restated on this repository's stack: the same cases (examples/1D_euler_convergence, 2D/3D_advection_convergence,
1D_sod_convergence via mfc_b200.cases, checked key by key against the reference's case files), the same error measures and
the same acceptance thresholds.

* h sweep in cell-shift mode (run_h_sweep, cell_shift = 1): T = h, compare q(T) with the initial field rolled by one cell
  along every axis; the L2 error then scales like h^(p+1), reported order = fitted slope - 1 >= expected - tol;
* dt sweep (run_dt_sweep): N = 512, one full period, vary the CFL number; slope in dt >= expected - tol;
* Sod L1 self-convergence (run_sod_l1): each N against 2N cell-averaged; rate >= expected - tol;
* conservation of every listed conserved variable to 1e-10 (CONS_TOL) in all advection runs.

CPU leg: the oracle (measured: WENO5 5.0, WENO3 1.5, WENO1 1.0, MUSCL2 2.0, TENO5 5.0 in 1-D and 2-D; RK1 1.0, RK2 2.0, RK3 2.95;
Sod L1 rates 0.61 - 1.01).  GPU leg (``-m gpu``): the product build through the C ABI, including the 3-D cases.
"""
import math

import numpy as np
import pytest

from mfc_b200 import cases, driver, pre_process
from mfc_b200.params import params_from_case

CONS_TOL = 1e-10  # convergence.py:43


def _run(case, backend):
    P = params_from_case(case)
    cbs, ccs = pre_process.generate_grid(case)
    _, cons = pre_process.initial_condition(case, ccs)
    if backend == "oracle":
        from oracle_lib import Oracle
        st = Oracle(P, cbs, omp=P.num_dims == 3)
    else:
        from mfc_b200.lib import Mfc5
        st = Mfc5(P, cbs)
    st.set_cons(cons)
    driver.run(st, case)
    out = st.get_cons()
    if backend != "oracle":
        st.close()
    return cons, out


def _slope(errors, xs):
    return float(np.polyfit(np.log(xs), np.log(errors), 1)[0])


def _conservation(q0, qT, idxs):
    return max(abs(float(np.sum(qT[i - 1])) - float(np.sum(q0[i - 1]))) / (abs(float(np.sum(q0[i - 1]))) + 1e-300) for i in idxs)


def h_sweep(backend, ndim, resolutions, expected, tol, **scheme):
    errors, hs = [], []
    cons_idx = {1: (1, 2, 3), 2: (1, 4), 3: (1, 5)}[ndim]  # cases.py:26-28
    for N in resolutions:
        h = 1.0 / N
        q0, qT = _run(cases.advection_convergence(ndim, N, t_end=1 * h, **scheme), backend)
        ref = np.roll(q0[0], shift=1, axis=tuple(range(q0[0].ndim - ndim, q0[0].ndim)))  # the wave moves by +1 cell per axis
        errors.append(math.sqrt(float(np.sum((qT[0] - ref) ** 2)) * h ** ndim))
        hs.append(h)
        assert _conservation(q0, qT, cons_idx) < CONS_TOL
    order = _slope(errors, hs) - 1.0
    assert order >= expected - tol, f"spatial order {order:.2f} < {expected - tol:.1f}; errors {errors}"
    return order


def dt_sweep(backend, cfls, expected, tol, time_stepper):
    errors, dts = [], []
    for cfl in cfls:
        case = cases.advection_convergence(1, 512, order=5, cfl=cfl, time_stepper=time_stepper)
        q0, qT = _run(case, backend)
        errors.append(math.sqrt(float(np.sum((qT[0] - q0[0]) ** 2)) / 512))
        dts.append(float(case["dt"]))
        assert _conservation(q0, qT, (1, 2, 3)) < CONS_TOL
    rate = _slope(errors, dts)
    assert rate >= expected - tol, f"temporal rate {rate:.2f} < {expected - tol:.1f}; errors {errors}"
    return rate


def sod_l1(backend, resolutions, expected, tol, **scheme):
    rho = [_run(cases.sod_convergence(N, **scheme), backend)[1][0].ravel() for N in resolutions]
    errors, hs = [], []
    for i in range(len(resolutions) - 1):
        avg = 0.5 * (rho[i + 1][0::2] + rho[i + 1][1::2])
        errors.append(float(np.sum(np.abs(rho[i] - avg))) / resolutions[i])
        hs.append(1.0 / resolutions[i])
    rate = _slope(errors, hs)
    assert rate >= expected - tol, f"L1 rate {rate:.2f} < {expected - tol:.1f}; errors {errors}"
    return rate


# (label, scheme arguments, expected order, tolerance, resolutions): cases.py:33-39, 47-53, 62-68
SCHEMES_1D = [
    ("WENO5", dict(order=5, cfl=0.02), 5, 0.3, [32, 64, 128]),
    ("WENO3", dict(order=3, cfl=0.02), 1.5, 0.3, [64, 128, 256]),
    ("WENO1", dict(order=1, cfl=0.02), 1, 0.2, [64, 128, 256]),
    ("MUSCL2", dict(muscl=True, muscl_lim=0, cfl=0.02), 2, 0.3, [64, 128, 256]),
    ("TENO5", dict(order=5, teno=True, teno_ct=1e-6, cfl=0.02), 5, 0.3, [32, 64, 128]),
]
SCHEMES_2D = [
    ("WENO5", dict(order=5, cfl=0.02), 5, 0.3, [32, 64, 96]),
    ("WENO3", dict(order=3, cfl=0.02), 1.5, 0.3, [32, 64, 128]),
    ("WENO1", dict(order=1, cfl=0.02), 1, 0.2, [32, 64, 128]),
    ("MUSCL2", dict(muscl=True, muscl_lim=0, cfl=0.02), 2, 0.3, [32, 64, 128]),
    ("TENO5", dict(order=5, teno=True, teno_ct=1e-6, cfl=0.02), 5, 0.3, [32, 64, 96]),
]
SCHEMES_3D = [(l, s, e, t, [32, 64]) for l, s, e, t, _ in SCHEMES_1D]
# cases.py:74-86 (WENO7 / TENO7: outside this path); min_N = 128 for SUPERBEE
SOD = [
    ("WENO1", dict(order=1), 1, 0.5, [128, 256, 512, 1024]),
    ("WENO3", dict(order=3), 1, 0.3, [128, 256, 512, 1024]),
    ("WENO5", dict(order=5), 1, 0.3, [128, 256, 512, 1024]),
    ("MUSCL-minmod", dict(muscl=True, muscl_lim=1), 1, 0.3, [128, 256, 512, 1024]),
    ("MUSCL-MC", dict(muscl=True, muscl_lim=2), 1, 0.3, [128, 256, 512, 1024]),
    ("MUSCL-VanLeer", dict(muscl=True, muscl_lim=4), 1, 0.3, [128, 256, 512, 1024]),
    ("MUSCL-SUPERBEE", dict(muscl=True, muscl_lim=5), 1, 0.5, [128, 256, 512, 1024]),
    ("TENO5", dict(order=5, teno=True, teno_ct=1e-6), 1, 0.3, [128, 256, 512, 1024]),
]
TEMPORAL = [("RK1", 1, 1, 0.1, [0.10, 0.05]), ("RK2", 2, 2, 0.2, [0.50, 0.25]), ("RK3", 3, 3, 0.3, [0.50, 0.25])]  # cases.py:90-94

_ids = lambda rows: [r[0] for r in rows]


# ---- CPU: the oracle -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("label,scheme,expected,tol,res", SCHEMES_1D, ids=_ids(SCHEMES_1D))
def test_oracle_spatial_order_1d(label, scheme, expected, tol, res):
    h_sweep("oracle", 1, res, expected, tol, **scheme)


@pytest.mark.parametrize("label,scheme,expected,tol,res", SCHEMES_2D, ids=_ids(SCHEMES_2D))
def test_oracle_spatial_order_2d(label, scheme, expected, tol, res):
    h_sweep("oracle", 2, res, expected, tol, **scheme)


@pytest.mark.parametrize("label,ts,expected,tol,cfls", TEMPORAL[1:], ids=_ids(TEMPORAL[1:]))  # RK1 (33 k steps): GPU leg only
def test_oracle_temporal_order(label, ts, expected, tol, cfls):
    dt_sweep("oracle", cfls, expected, tol, ts)


@pytest.mark.parametrize("label,scheme,expected,tol,res", SOD, ids=_ids(SOD))
def test_oracle_sod_l1(label, scheme, expected, tol, res):
    sod_l1("oracle", res, expected, tol, **scheme)


# ---- GPU: the product build through the C ABI -------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("label,scheme,expected,tol,res", SCHEMES_1D, ids=_ids(SCHEMES_1D))
def test_gpu_spatial_order_1d(label, scheme, expected, tol, res):
    h_sweep("gpu", 1, res, expected, tol, **scheme)


@pytest.mark.gpu
@pytest.mark.parametrize("label,scheme,expected,tol,res", SCHEMES_2D, ids=_ids(SCHEMES_2D))
def test_gpu_spatial_order_2d(label, scheme, expected, tol, res):
    h_sweep("gpu", 2, res, expected, tol, **scheme)


@pytest.mark.gpu
@pytest.mark.parametrize("label,scheme,expected,tol,res", SCHEMES_3D, ids=_ids(SCHEMES_3D))
def test_gpu_spatial_order_3d(label, scheme, expected, tol, res):
    h_sweep("gpu", 3, res, expected, tol, **scheme)


@pytest.mark.gpu
@pytest.mark.parametrize("label,ts,expected,tol,cfls", TEMPORAL, ids=_ids(TEMPORAL))
def test_gpu_temporal_order(label, ts, expected, tol, cfls):
    dt_sweep("gpu", cfls, expected, tol, ts)


@pytest.mark.gpu
@pytest.mark.parametrize("label,scheme,expected,tol,res", SOD, ids=_ids(SOD))
def test_gpu_sod_l1(label, scheme, expected, tol, res):
    sod_l1("gpu", res, expected, tol, **scheme)
