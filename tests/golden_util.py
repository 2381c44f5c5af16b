"""Shared helpers for the golden-file parity tests (the reference's own regression data)."""
import json
import os

import numpy as np

from mfc_b200 import pre_process
from mfc_b200.params import params_from_case

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

with open(os.path.join(GOLD, "cases.json")) as _f:
    _ALL = json.load(_f)
# MTHINC (int_comp = 2) exists in the CPU oracle only so far (tests/test_oracle_next.py); every other list of cases is the product's
ORACLE_ONLY = {u: c for u, c in _ALL.items() if str(c["params"].get("int_comp", 0)) in ("2", "mthinc")}
VISCOUS = sorted(u for u, c in _ALL.items() if str(c["params"].get("viscous", "F")) == "T" and u not in ORACLE_ONLY)
ORDER7 = sorted(u for u, c in _ALL.items() if int(c["params"].get("weno_order", 0) or 0) == 7)
CASES = {u: c for u, c in _ALL.items() if u not in ORACLE_ONLY}


def golden(uuid):
    return dict(np.load(os.path.join(GOLD, uuid + ".npz")))


def golden_field(g, kind, var, rank, step, shape):
    """Golden line -> array (nz, ny, nx).  File order is x outer, y middle, z inner
    (reference src/simulation/m_data_output.fpp:561-569)."""
    a = g[f"{kind}.{var}.{rank:02d}.{step:06d}"]
    nz, ny, nx = shape
    return a.reshape(nx, ny, nz).transpose(2, 1, 0)


def compare(candidate, gold, tol):
    """The reference's rule, toolchain/mfc/packer/tol.py:10-20 / errors.py:15-25:
    pass iff abs_err <= tol or rel_err <= tol; NaN fails.  Returns (ok, max_abs, max_rel_of_failing)."""
    c = np.asarray(candidate, dtype=np.float64).ravel()
    g = np.asarray(gold, dtype=np.float64).ravel()
    if np.isnan(c).any() or np.isnan(g).any():
        return False, np.nan, np.nan
    abs_err = np.abs(c - g)
    denom = np.maximum(np.abs(g), np.finfo(float).tiny)
    rel_err = abs_err / denom
    bad = (abs_err > tol) & (rel_err > tol)
    worst = float(np.max(np.minimum(abs_err, rel_err))) if c.size else 0.0
    return (not bad.any()), float(abs_err.max()), worst


def setup_case(uuid, **kw):
    """(case dict, params, cell boundaries, initial conservative state, Dirichlet buffers) for a single-rank case."""
    case = CASES[uuid]["params"]
    P = params_from_case(case, **kw)
    cbs, ccs = pre_process.generate_grid(case)
    prim, cons = pre_process.initial_condition(case, ccs)
    return case, P, cbs, cons, pre_process.dirichlet_buffers(P, prim)
