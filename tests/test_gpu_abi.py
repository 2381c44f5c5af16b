"""The drop-in boundary as the Fortran host would use it.

* ``ghosted = 1``: the buffered ``q_cons_vf(i)%sf(-b:m+b, -b:n+b, -b:p+b)`` layout on upload and download
  (reference ``src/simulation/m_time_steppers.fpp:179-185``, ``src/common/m_derived_types.fpp:21-24``);
* ``mfc5_download_prim(ghosted = 1)``: the ghost-filled ``q_prim_vf`` contract of ``m_rhs.fpp:884-898`` --
  edges and corners included, filled x -> y -> z as ``m_boundary_common.fpp:99-113,143-149``;
* a plain-C program (``tests/c_driver/abi_driver.c``, gcc against ``include/mfc5.h``) that performs the shim's call
  sequence create -> set_grid -> finalize_setup -> upload(ghosted = 1) -> tvd_rk -> download;
* the CUDA-graph cache of ``mfc5_tvd_rk_n`` across ``mfc5_reduce`` / ``mfc5_compute_rhs`` / ``mfc5_set_dirichlet``
  (a graph captured right after a reduction must still contain its cons -> prim pass).
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from golden_util import CASES, setup_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# 3-D two-fluid (-3), 1-D, 2-D GRCBC / Dirichlet / slip wall, 3-D slip (-15) / no-slip (-16) / Dirichlet (-17) walls, the 3-D
# shock-droplet example (-6 / -2 / -3 mixed: reflective sign flips on edges and corners), 3-D characteristic (-6)
GHOST_SET = ["A0B82851", "2AB32975", "EB6687AB", "10DBA76B", "4D7926CD", "6249EDDE", "8BCBAED8", "9DCBE54F", "B5F0C655", "5010B814"]
GHOST_SET = [u for u in GHOST_SET if u in CASES]


def _pair(uuid, strict=True):
    from mfc_b200.lib import Mfc5
    from oracle_lib import Oracle
    case, P, cbs, cons, dbuf = setup_case(uuid)
    g = Mfc5(P, cbs, dirichlet=dbuf, strict=strict)
    o = Oracle(P, cbs, omp=True)
    for (d, side), buf in dbuf.items():
        o.set_dirichlet(d, side, buf)
    return case, P, cbs, cons, dbuf, g, o


def _embed(g, cons, fill=np.nan):
    """Interior-packed state -> buffered boxes whose ghost cells hold `fill` (the library must never read them)."""
    box = np.full(g.box_shape, fill)
    b, P = g.b, g.P
    sl = (slice(None), slice(b, b + P.p + 1) if P.p > 0 else slice(0, 1), slice(b, b + P.n + 1) if P.n > 0 else slice(0, 1),
          slice(b, b + P.m + 1))
    box[sl] = cons
    return box, sl


pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("uuid", GHOST_SET)
def test_ghosted_upload_download_roundtrip_and_rhs(uuid):
    """upload(ghosted = 1) reads the interior only; download(ghosted = 1) writes the interior back bit for bit; the RHS of
    the state uploaded through the buffered layout equals the oracle's."""
    case, P, cbs, cons, dbuf, g, o = _pair(uuid)
    box, sl = _embed(g, cons)
    g.set_cons_box(box)
    back = g.get_box("cons")
    np.testing.assert_array_equal(back[sl], cons)
    np.testing.assert_array_equal(g.get_cons(), cons)   # and the same state through the interior layout
    o.set_cons(cons)
    g.compute_rhs()
    o.compute_rhs()
    r_o = o.get_rhs()
    r_g = g.get_box("rhs")[sl]
    assert np.max(np.abs(r_g - r_o)) <= 1e-13 * (np.max(np.abs(r_o)) + 1e-300)


@pytest.mark.parametrize("uuid", GHOST_SET)
def test_ghosted_prim_matches_oracle_box(uuid):
    """q_prim_vf with every ghost cell (faces, edges, corners) as the reference's sequential x -> y -> z fill leaves them:
    strict build vs the oracle's buffered box, bit for bit."""
    case, P, cbs, cons, dbuf, g, o = _pair(uuid)
    box, sl = _embed(g, cons)
    g.set_cons_box(box)
    o.set_cons(cons)
    g.compute_rhs()
    o.compute_rhs()
    a, b = g.get_box("prim"), o.get_box("prim")
    assert a.shape == b.shape
    np.testing.assert_array_equal(a, b)
    # after a step the library converts the new state itself before handing out q_prim_vf
    dt = float(case["dt"])
    g.step(dt)
    o.step(dt)
    o.compute_rhs()   # refreshes the oracle's primitive box for the new state
    np.testing.assert_array_equal(g.get_box("prim"), o.get_box("prim"))


@pytest.mark.parametrize("uuid", ["A0B82851", "EB6687AB"])
def test_ghosted_state_after_steps(uuid):
    """s_tvd_rk through the buffered layout: 5 steps, strict build == oracle bit for bit; the ghost cells of the downloaded
    q_cons_vf are the device's own (zeros), never NaN / garbage."""
    case, P, cbs, cons, dbuf, g, o = _pair(uuid)
    box, sl = _embed(g, cons, fill=-7.25)
    g.set_cons_box(box)
    o.set_cons(cons)
    dt = float(case["dt"])
    for _ in range(5):
        g.step(dt)
        o.step(dt)
    out = np.full(g.box_shape, -7.25)
    g.get_box("cons", out=out)
    np.testing.assert_array_equal(out[sl], o.get_cons())
    ghost = np.ones(out.shape, dtype=bool)
    ghost[sl] = False
    assert np.all(out[ghost] == 0.0)


def _build_driver(tmp_path):
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.join(ROOT, "mfc_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_driver", "abi_driver.c"),
                           "-L", libdir, "-lmfc5", f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


@pytest.mark.parametrize("use_n", [0, 1])
@pytest.mark.parametrize("uuid", ["A0B82851", "04E1CB26"])
def test_c_program_drives_the_abi_like_the_fortran_shim(uuid, use_n, tmp_path):
    """gcc-compiled C host: create -> set_grid -> finalize_setup -> upload(ghosted = 1) -> tvd_rk x 6 -> reduce -> download
    cons and prim (ghosted = 1), product build, checked against the oracle at the reference's golden rule."""
    from golden_util import compare
    from oracle_lib import Oracle
    case, P, cbs, cons, dbuf = setup_case(uuid)
    o = Oracle(P, cbs, omp=True)
    exe = _build_driver(tmp_path)
    nsteps, dt = 6, float(case["dt"])
    b = o.b
    box = np.full(o.box_shape, np.nan)
    sl = (slice(None), slice(b, b + P.p + 1) if P.p > 0 else slice(0, 1), slice(b, b + P.n + 1) if P.n > 0 else slice(0, 1),
          slice(b, b + P.m + 1))
    box[sl] = cons
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(bytes(P))
        f.write(struct.pack("iid", nsteps, use_n, dt))
        for d in range(P.num_dims):
            cb = np.ascontiguousarray(cbs[d], dtype=np.float64)
            f.write(struct.pack("i", cb.size))
            f.write(cb.tobytes())
        f.write(box.tobytes())
    out = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "kernel launches" in out.stdout
    raw = np.fromfile(fout, dtype=np.float64)
    n = box.size
    q = raw[:n].reshape(box.shape)
    pr = raw[n:2 * n].reshape(box.shape)
    icfl, tavg = raw[2 * n], raw[2 * n + 1]
    o.set_cons(cons)
    for _ in range(nsteps):
        o.step(dt)
    ref = o.get_cons()
    for v in range(ref.shape[0]):
        ok, mabs, worst = compare(q[sl][v], ref[v], 1e-12)
        assert ok, f"var {v + 1}: {worst:.3e}"
    assert np.isfinite(q).all()   # the NaN the caller left in its ghost cells was never read, and is overwritten by zeros
    o.compute_rhs()
    pb = o.get_box("prim")
    for v in range(pb.shape[0]):   # abs-or-rel like the reference's packer: the transverse velocities of this case are 1e-18 noise
        ok, mabs, worst = compare(pr[v], pb[v], 1e-10)
        assert ok, f"q_prim_vf var {v + 1} (ghost cells included): {worst:.3e}"
    assert abs(icfl - o.icfl_max(dt)) <= 1e-9 * o.icfl_max(dt)
    assert use_n == 1 or tavg > 0.0


def test_graph_captured_after_reduce_keeps_its_cons2prim_pass():
    """steps(2); reduce; steps(2); steps(2) -- the graph for the parity reached after the reduction is captured
    with prim_valid = true; it must still start with cons -> prim.  Strict build: bit-equal to six single steps."""
    from mfc_b200.lib import Mfc5
    case, P, cbs, cons, dbuf = setup_case("A0B82851")
    dt = float(case["dt"])
    a = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    b = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    a.set_cons(cons)
    b.set_cons(cons)
    for _ in range(7):
        a.step(dt)
    b.steps(dt, 1)          # eager warm-up step
    b.steps(dt, 2)          # captures key (cur = 1), replays key 0
    assert b.reduce(2) == 0.0
    b.steps(dt, 2)
    b.compute_rhs()         # leaves prim_valid = true as well
    b.steps(dt, 2)
    b.sync()
    assert b.graph_state == 1
    np.testing.assert_array_equal(a.get_cons(), b.get_cons())


def test_odd_step_counts_after_reduce_and_rhs():
    """Every (parity, prim_valid) combination at capture time: RK3 flips the ping-pong parity each step."""
    from mfc_b200.lib import Mfc5
    case, P, cbs, cons, dbuf = setup_case("04E1CB26")
    dt = float(case["dt"])
    a = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    b = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    a.set_cons(cons)
    b.set_cons(cons)
    n = 0
    for k, pre in ((1, None), (1, "reduce"), (1, "rhs"), (2, "reduce"), (3, "rhs"), (1, "prim"), (2, None)):
        if pre == "reduce":
            b.icfl_max(dt)
        elif pre == "rhs":
            b.compute_rhs()
        elif pre == "prim":
            b.get_prim()
        b.steps(dt, k)
        n += k
    for _ in range(n):
        a.step(dt)
    np.testing.assert_array_equal(a.get_cons(), b.get_cons())


def test_set_dirichlet_after_captured_steps():
    """mfc5_set_dirichlet after graphs exist: same allocation is reused (or the graphs are re-captured); new face values
    take effect and nothing reads freed memory."""
    from mfc_b200.lib import Mfc5
    uuid = next((u for u in CASES if CASES[u]["ppn"] == 1 and any(int(CASES[u]["params"].get(f"bc_{x}%{e}", 0) or 0) == -17
                                                                     for x in "xyz" for e in ("beg", "end"))), None)
    if uuid is None:
        pytest.skip("no Dirichlet golden")
    case, P, cbs, cons, dbuf = setup_case(uuid)
    assert dbuf
    dt = float(case["dt"])
    a = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    b = Mfc5(P, cbs, dirichlet=dbuf, strict=True)
    a.set_cons(cons)
    b.set_cons(cons)
    b.steps(dt, 3)
    b.sync()
    for (d, side), buf in dbuf.items():   # same values again, through the entry point, after the graphs were captured
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        b._ck(b.L.mfc5_set_dirichlet(b.h, d, side, buf.ctypes.data_as(C.POINTER(C.c_double))))
    b.steps(dt, 3)
    for _ in range(6):
        a.step(dt)
    np.testing.assert_array_equal(a.get_cons(), b.get_cons())
