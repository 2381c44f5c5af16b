"""Direction-permutation symmetry (a free invariant of the scheme).

On a cube whose grid and boundary conditions are the same in x, y and z, rotating the initial state (x -> y -> z -> x, velocity
components permuted with the axes) must rotate the result.  The three sweeps are three different kernels on the GPU (x: warp
shuffles along the row; y, z: marching pencils, z with the fused RK update), with per-direction WENO tables and boundary code,
so this checks them against each other; the only admissible difference is the summation order of the three directional
contributions to the right-hand side (a few ulp).  Cases: reference goldens' parameter sets (cubic, extrapolation boundaries)
and the BASELINE workload examples/3D_performance_test at 40^3 (stretched grid, reflective / characteristic boundaries in every
direction, case-optimised kernels on the GPU).
"""
import numpy as np
import pytest

from golden_util import CASES
from mfc_b200 import cases, pre_process
from mfc_b200.params import params_from_case


def _cases():
    return {
        "A0B82851": (dict(CASES["A0B82851"]["params"]), 10),   # 2 fluids, HLLC, mpp_lim, WENO5-JS
        "04E1CB26": (dict(CASES["04E1CB26"]["params"]), 10),   # 1 fluid, WENO5-M
        "758D0268": (dict(CASES["758D0268"]["params"]), 10),   # 2 fluids, HLL
        "C5_perf": (cases.perf_test_3d(m=39, steps=5), 5),     # BASELINE C5: stretched, BC -2 / -6, WENO5-M + HLLC
    }


def rotate(q, nf):
    """(N, nz, ny, nx) -> the state seen with new x = old y, new y = old z, new z = old x"""
    r = np.ascontiguousarray(q.transpose(0, 3, 1, 2))   # r[v, a, b, c] = q[v, b, c, a]
    mom = r[nf:nf + 3].copy()
    r[nf], r[nf + 1], r[nf + 2] = mom[1], mom[2], mom[0]  # u' = v, v' = w, w' = u
    return r


def _run(name, backend, **kw):
    case, nsteps = _cases()[name]
    P = params_from_case(case, **{k: v for k, v in kw.items() if k == "fast_math"})
    cbs, ccs = pre_process.generate_grid(case)
    for d in (1, 2):
        np.testing.assert_array_equal(cbs[d], cbs[0])  # the premise: the same grid in every direction
    _, cons = pre_process.initial_condition(case, ccs)
    # make the initial state itself asymmetric (the test cases' patches are symmetric under the rotation)
    nf = P.num_fluids
    nz, ny, nx = cons.shape[1:]
    z, y, x = np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij")
    bump = 1.0 + 0.05 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * (y + 0.3)) + 0.03 * np.sin(2 * np.pi * (z + 0.1))
    cons = cons.copy()
    rho = cons[:nf].sum(axis=0)
    vel = (0.1 * bump, -0.07 * bump * bump, 0.05 / bump)
    cons[nf + 3] += 0.5 * rho * sum(v * v for v in vel) - 0.5 * sum(cons[nf + d] ** 2 for d in range(3)) / rho
    for d in range(3):
        cons[nf + d] = rho * vel[d]
    outs = []
    for q0 in (cons, rotate(cons, nf)):
        if backend == "oracle":
            from oracle_lib import Oracle
            st = Oracle(P, cbs, omp=True)
        else:
            from mfc_b200.lib import Mfc5
            st = Mfc5(P, cbs, strict=kw.get("strict", False))
        st.set_cons(q0)
        dt = float(case["dt"])
        for _ in range(nsteps):
            st.step(dt)
        outs.append(st.get_cons())
        if backend != "oracle":
            st.close()
    a, b = rotate(outs[0], nf), outs[1]
    assert not np.array_equal(outs[0], cons)
    worst = 0.0
    for v in range(a.shape[0]):
        worst = max(worst, float(np.max(np.abs(a[v] - b[v])) / (np.max(np.abs(b[v])) + 1e-300)))
    return worst


@pytest.mark.parametrize("name", ["A0B82851", "04E1CB26", "758D0268", "C5_perf"])
def test_oracle_permutation_symmetry(name):
    assert _run(name, "oracle") <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "fast", "ref_order"])
@pytest.mark.parametrize("name", ["A0B82851", "04E1CB26", "758D0268", "C5_perf"])
def test_gpu_permutation_symmetry(name, mode):
    worst = _run(name, "gpu", strict=(mode == "strict"), fast_math=1 if mode == "fast" else 0)
    assert worst <= 1e-13, f"{name} {mode}: {worst:.3e}"
