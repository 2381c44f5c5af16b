"""f-1 / f-2: case input, on-disk formats and the p_main time loop.

The loop is driven here with the CPU oracle as the state object (test infrastructure; the product's
``make_state`` default is the CUDA library and has no fallback); the GPU version of the same checks is
``tests/test_gpu_driver.py``.  Reference data: the ``D/*.dat`` lines of the reference's own goldens, compared
with the packer's rule after a round trip through the formatted files.
"""
import json
import os
import sys

import numpy as np
import pytest

from golden_util import CASES, compare, golden
from mfc_b200 import case_io, data_io, main
from oracle_lib import Oracle


class OracleState(Oracle):
    """The oracle's ICFL uses the primitive variables of its last RHS evaluation; the driver asks for the ICFL of the
    current state (what the CUDA library's reduction does), so convert first."""

    def icfl_max(self, dt):
        self.cons2prim()
        return super().icfl_max(dt)


def oracle_state(P, cbs, gws, dirichlet):
    o = OracleState(P, cbs, ghost_widths=gws)
    for (d, side), buf in (dirichlet or {}).items():
        o.set_dirichlet(d, side, buf)
    return o


class SerialWorld:
    rank, size, local_rank, dist = 0, 1, 0, None

    def bcast(self, x):
        return x

    def max(self, x):
        return x

    def barrier(self):
        pass


def run_case(case, case_dir, targets=("pre_process", "simulation")):
    lines = []
    w = SerialWorld()
    if "pre_process" in targets:
        main.run_pre_process(case, case_dir, w, log=lines.append)
    out = None
    if "simulation" in targets:
        out = main.run_simulation(case, case_dir, w, make_state=oracle_state, log=lines.append)
    return out, lines


def test_namelist_round_trip(tmp_path):
    case = {"m": 99, "dt": 1e-06, "weno_eps": 1e-16, "mapped_weno": "T", "case_dir": "'.'", "bc_x%beg": -3, "fluid_pp(1)%gamma": 2.5,
            "patch_icpp(1)%alter_patch(1)": "F", "x_domain%end": 1.0, "file_name": "long string"}
    p = tmp_path / "simulation.inp"
    case_io.write_namelist(str(p), case)
    text = p.read_text()
    assert text.startswith("&user_inputs\n") and text.endswith("&end/\n")
    assert "mapped_weno = T\n" in text and "file_name = 'long string'\n" in text  # toolchain/mfc/case.py:100-103
    back = case_io.read_namelist(str(p))
    for k, v in case.items():
        if k == "case_dir":
            continue
        assert back[k] == v, k
    (tmp_path / "f.inp").write_text("&user_inputs\nm = 10\ndt = 1.5d-3\nmpp_lim = .true.\n&end/\n")
    b = case_io.read_namelist(str(tmp_path / "f.inp"))
    assert b == {"m": 10, "dt": 1.5e-3, "mpp_lim": "T"}


def test_load_case_py(tmp_path):
    """A case file is executed with ``--mfc <json>`` and extra arguments, like toolchain/mfc/run/input.py:188."""
    src = ("import json, argparse\np = argparse.ArgumentParser()\np.add_argument('--mfc', type=json.loads, default={})\n"
           "p.add_argument('--scale', type=int, default=1)\na = p.parse_args()\nprint(json.dumps({'m': 10 * a.scale, 'dt': 1e-3}))\n")
    f = tmp_path / "case.py"
    f.write_text(src)
    assert case_io.load_case(str(f)) == {"m": 10, "dt": 1e-3}
    assert case_io.load_case(str(f), ["--scale", "3"])["m"] == 30
    (tmp_path / "bad.py").write_text("raise SystemExit(3)\n")
    with pytest.raises(case_io.CaseError):
        case_io.load_case(str(tmp_path / "bad.py"))
    with pytest.raises(case_io.CaseError):
        case_io.load_case(str(tmp_path / "missing.py"))


def test_fortran_records_and_text(tmp_path):
    rng = np.random.default_rng(7)
    cons = rng.standard_normal((5, 3, 4, 6)) * 1e3
    cbs = [np.linspace(0, 1, 7), np.linspace(-1, 1, 5), np.linspace(0, 2, 4)]
    data_io.write_serial_restart(str(tmp_path), 0, 7, cbs, cons)
    raw = open(tmp_path / "p_all" / "p0" / "7" / "q_cons_vf2.dat", "rb").read()
    assert len(raw) == 4 + 8 * 72 + 4 and int.from_bytes(raw[:4], "little") == 576 == int.from_bytes(raw[-4:], "little")
    cb2, c2 = data_io.read_serial_restart(str(tmp_path), 0, 7, 5, 3, 2, 5)
    assert np.array_equal(c2, cons) and all(np.array_equal(a, b) for a, b in zip(cb2, cbs))
    with pytest.raises(IOError):
        data_io.read_serial_restart(str(tmp_path), 0, 8, 5, 3, 2, 5)
    data_io.write_D(str(tmp_path), 0, 7, cbs, cons)
    lines = open(tmp_path / "D" / "cons.1.00.000007.dat").read().split("\n")
    assert len(lines[0]) == 160 and lines[3] == ""  # 4F40.14, blank line after the z loop (nz = 3)
    packed = data_io.pack_D(str(tmp_path))
    assert sorted(packed) == [f"D/cons.{i}.00.000007.dat" for i in range(1, 6)]
    for i in range(5):
        want = cons[i].transpose(2, 1, 0).ravel()  # x outer, z inner
        assert np.max(np.abs(packed[f"D/cons.{i + 1}.00.000007.dat"] - want)) <= 0.5e-14 + np.spacing(4.0e3)


def test_parallel_io_blocks(tmp_path):
    rng = np.random.default_rng(3)
    glb = rng.standard_normal((4, 5, 6, 9))
    with pytest.raises(IOError):  # the file is created collectively first
        data_io.write_parallel_restart(str(tmp_path), 12, (8, 5, 4), (0, 0, 0), glb[:, :, :, :5])
    data_io.create_parallel_restart(str(tmp_path), 12, (8, 5, 4), 4)
    data_io.write_parallel_restart(str(tmp_path), 12, (8, 5, 4), (0, 0, 0), glb[:, :, :, :5])
    data_io.write_parallel_restart(str(tmp_path), 12, (8, 5, 4), (5, 0, 0), glb[:, :, :, 5:])
    assert os.path.getsize(tmp_path / "restart_data" / "lustre_12.dat") == glb.size * 8
    assert np.array_equal(np.fromfile(tmp_path / "restart_data" / "lustre_12.dat").reshape(glb.shape), glb)
    back = data_io.read_parallel_restart(str(tmp_path), 12, (8, 5, 4), (5, 0, 0), (3, 5, 4), 4)
    assert np.array_equal(back, glb[:, :, :, 5:])
    cbs = [np.linspace(0, 1, 10), np.linspace(0, 1, 7), np.linspace(0, 1, 6)]
    data_io.write_parallel_grid(str(tmp_path), cbs)
    assert all(np.array_equal(a, b) for a, b in zip(data_io.read_parallel_grid(str(tmp_path), 8, 5, 4), cbs))


@pytest.mark.parametrize("uuid", ["2AB32975", "BCA926AC", "05D092C5", "A0B82851"])
def test_case_through_files_matches_golden(uuid, tmp_path):
    """pre_process + simulation through the on-disk formats; every golden D/ entry with the packer's tolerance rule."""
    case = dict(CASES[uuid]["params"])
    out, lines = run_case(case, str(tmp_path))
    packed = data_io.pack_D(str(tmp_path))
    g = golden(uuid)
    tol = float(CASES[uuid]["tol"])
    assert len(g) > 0
    for key, want in g.items():
        name = f"D/{key}.dat"
        assert name in packed, name
        ok, amax, worst = compare(packed[name], want, max(tol, 1e-12))
        assert ok, (uuid, key, amax, worst)
    assert any("Performance:" in l and "ns/gp/eq/rhs" in l for l in lines)
    assert any(l.startswith(" [  0%]  Time step        1 of") for l in lines)
    rows = open(tmp_path / "time_data.dat").read().split("\n")
    assert rows[0] == "     Ranks         s/step   ns/gp/eq/rhs" and len(rows[1]) == 40


def test_restart_round_trip(tmp_path):
    """0 -> 50 in one run vs 0 -> 25, restart from the p_all files, 25 -> 50 (reference tests 1A379909 / C3A4DBAE do the same).
    Not bitwise, in the reference either: the last step of the first leg runs with dt = finaltime - mytime (m_start_up.fpp:589-594),
    which differs from dt in the last bits."""
    case = dict(CASES["6B4B738B"]["params"]) if "6B4B738B" in CASES else dict(CASES["2AB32975"]["params"])
    a, _ = run_case(dict(case), str(tmp_path / "a"))
    c1 = dict(case, t_step_stop=25, t_step_save=25)
    run_case(c1, str(tmp_path / "b"))
    c2 = dict(case, t_step_start=25, t_step_stop=50, t_step_save=25)
    b, _ = run_case(c2, str(tmp_path / "b"), targets=("simulation",))
    ok, amax, worst = compare(b["cons"], a["cons"], 1e-12)
    assert ok, (amax, worst)
    assert os.path.isdir(tmp_path / "b" / "p_all" / "p0" / "50")


def test_run_time_info(tmp_path):
    case = dict(CASES["2AB32975"]["params"], run_time_info="T", t_step_stop=5, t_step_save=5)
    run_case(case, str(tmp_path))
    txt = open(tmp_path / "run_time.inf").read().split("\n")
    assert txt[0].startswith("Description: Stability information at each time-step")
    hdr = next(i for i, l in enumerate(txt) if "Time-step" in l)
    rows = [l for l in txt[hdr + 1:hdr + 6]]
    assert len(rows[0]) == 4 * 13 + 9 + 30 and int(rows[4].split()[0]) == 4
    icfl = float(rows[0].split()[3])
    assert 0.0 < icfl < 1.0
    assert any(l.startswith("ICFL Max:") for l in txt)


def test_cli_parallel_io_round_trip(tmp_path):
    """pre_process via the CLI with parallel_io = T writes restart_data/ in the reference's layout."""
    case = dict(CASES["2AB32975"]["params"], parallel_io="T")
    f = tmp_path / "case.json"
    f.write_text(json.dumps(case))
    rc = main.main(["run", str(f), "-t", "pre_process"])
    assert rc == 0
    assert os.path.exists(tmp_path / "simulation.inp")
    m = int(case["m"])
    assert os.path.getsize(tmp_path / "restart_data" / "lustre_x_cb.dat") == (m + 2) * 8
    nsys = 2 * int(case["num_fluids"]) + 2
    assert os.path.getsize(tmp_path / "restart_data" / "lustre_0.dat") == nsys * (m + 1) * 8
    out, _ = run_case(case, str(tmp_path), targets=("simulation",))
    ref, _ = run_case(dict(case, parallel_io="F"), str(tmp_path / "serial"))
    assert np.array_equal(out["cons"], ref["cons"])


class OracleCfl(OracleState):
    """dt_min of the current state for the CFL-driven loop (the CUDA library: mfc5_reduce(what = 1)); numpy here,
    following s_compute_enthalpy / s_compute_speed_of_sound / s_compute_dt_from_cfl (m_sim_helpers.fpp:45-105,182-215)."""

    def reduce(self, what, dt=0.0):
        self.cons2prim()                 # what == 2 (NaN count) is only used by the driver to refresh the primitives
        if what != 1:
            return 0.0
        q = self.get_prim()
        P = self.P
        nf, nd = P.num_fluids, P.num_dims
        arho = [q[f].copy() for f in range(nf)]
        alf = [q[nf + nd + 1 + f].copy() for f in range(nf)]
        if P.mpp_lim:
            for _ in range(2):
                s = np.zeros_like(alf[0])
                for f in range(nf):
                    arho[f] = np.maximum(0.0, arho[f])
                    alf[f] = np.minimum(np.maximum(0.0, alf[f]), 1.0)
                    s = s + alf[f]
                for f in range(nf):
                    alf[f] = alf[f] / np.maximum(s, 1e-16)
        rho = sum(arho)
        g = sum(alf[f] * P.fluid_gamma[f] for f in range(nf))
        pinf = sum(alf[f] * P.fluid_pi_inf[f] for f in range(nf))
        vel = [q[nf + a] for a in range(nd)]
        vs = sum(v * v for v in vel)
        p = q[nf + nd]
        E = g * p + pinf + 0.5 * rho * vs
        H = (E + p) / rho
        c = np.sqrt((H - 0.5 * vs) / g)
        dts = [self._ds[a] / (np.abs(vel[a]) + c) for a in range(nd)]
        dtl = np.minimum.reduce(dts)
        if P.viscous:   # viscous limit, s_compute_dt_from_cfl m_sim_helpers.fpp:214-227 with the mixture Reynolds numbers of
            # s_convert_species_to_mixture_variables_kernel (m_variables_conversion.fpp:256-268)
            inv = []
            for i in range(2):
                idx = [f for f in range(nf) if P.fluid_Re[f][i] > 0]
                if idx:
                    Re = 1.0 / np.maximum(sum(alf[f] / P.fluid_Re[f][i] for f in idx), 1e-16)
                else:
                    Re = np.full_like(rho, -1.0e6)
                inv.append(1.0 / (rho * Re) if nd != 2 else (1.0 / Re) / rho)
            dsm = np.minimum.reduce([self._ds[a] + 0.0 * rho for a in range(nd)])
            dtl = np.minimum(dtl, dsm ** 2.0 / np.maximum(inv[0], inv[1]))
        return float(np.min(dtl))


def _make_cfl_oracle(P, cbs, gws, dirichlet):
    o = OracleCfl(P, cbs, ghost_widths=gws)
    shapes = [(1, 1, -1), (1, -1, 1), (-1, 1, 1)]
    o._ds = [np.diff(cbs[a]).reshape(shapes[a]) for a in range(P.num_dims)]
    return o


def test_cfl_driven_time_stepping(tmp_path):
    """cfl_adap_dt: dt = cfl_target * min(dx / (|u| + c)) every step, run bounded by t_stop, saves every t_save
    (s_compute_dt m_time_steppers.fpp:639-710, p_main.fpp:48-81)."""
    case = dict(CASES["2AB32975"]["params"], cfl_adap_dt="T", cfl_target=0.4, t_stop=0.02, t_save=0.01, n_start=0, run_time_info="T")
    lines = []
    w = SerialWorld()
    main.run_pre_process(case, str(tmp_path), w, log=lines.append)

    out = main.run_simulation(case, str(tmp_path), w, make_state=_make_cfl_oracle, log=lines.append)
    rows = [l for l in open(tmp_path / "run_time.inf").read().split("\n") if l.strip() and l.split()[0].isdigit()]
    dts = np.array([float(r.split()[1]) for r in rows])
    icfl = np.array([float(r.split()[3]) for r in rows])
    assert len(rows) == out["t_step"] and out["t_step"] > 10
    assert np.all(np.abs(icfl[:-1] - 0.4) < 2e-6)                # ICFL == cfl_target by construction (the last step is clipped)
    assert abs(dts.sum() - 0.02) < 1e-5
    assert os.path.isdir(tmp_path / "p_all" / "p0" / "1") and os.path.isdir(tmp_path / "p_all" / "p0" / "2")
    assert any(l.startswith(" [  0%] Time ") for l in lines)


@pytest.mark.parametrize("uuid", ["A078904B", "B54BB9D8", "AE02324F", "121D4ECA", "70EC99CE"])
def test_cfl_driven_goldens(uuid, tmp_path):
    """Reference goldens of CFL-driven runs through the driver loop: A078904B (2-D Kelvin-Helmholtz example, HLL, hard-coded
    profile 207; every D/ file of every save, selected by simulated time) and B54BB9D8 (3-D body forces with cfl_adap_dt:
    the acceleration follows the accumulated time)."""
    case = dict(CASES[uuid]["params"])
    w = SerialWorld()
    main.run_pre_process(case, str(tmp_path), w, log=lambda s: None)
    main.run_simulation(case, str(tmp_path), w, make_state=_make_cfl_oracle, log=lambda s: None)
    packed = data_io.pack_D(str(tmp_path))
    g = golden(uuid)
    assert len(g) >= (100 if uuid == "A078904B" else 10)
    for key, want in g.items():
        name = f"D/{key}.dat"
        assert name in packed, name
        ok, amax, worst = compare(packed[name], want, float(CASES[uuid]["tol"]))
        assert ok, (key, amax, worst)
