"""f-1 / f-2 on the GPU: ``python -m mfc_b200 run`` (pre_process + simulation, CUDA hot path) on the reference's own
regression cases, compared through the formatted ``D/`` files with the reference's goldens and packer rule."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from golden_util import CASES, compare, golden
from mfc_b200 import data_io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_cli(case, case_dir, nranks=1, extra=()):
    os.makedirs(case_dir, exist_ok=True)
    f = os.path.join(case_dir, "case.json")
    with open(f, "w") as fh:
        json.dump(case, fh)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if nranks == 1:
        cmd = [sys.executable, "-m", "mfc_b200", "run", f, *extra]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", "-m", "mfc_b200", "run", f, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def _check_golden(uuid, case_dir):
    packed = data_io.pack_D(case_dir)
    g = golden(uuid)
    tol = max(float(CASES[uuid]["tol"]), 1e-12)
    for key, want in g.items():
        name = f"D/{key}.dat"
        assert name in packed, name
        ok, amax, worst = compare(packed[name], want, tol)
        assert ok, (uuid, key, amax, worst)


@pytest.mark.parametrize("uuid", ["2AB32975", "BCA926AC", "05D092C5", "A0B82851", "04E1CB26", "B49877F3", "A078904B", "C9E0B721", "8370A3D9", "10041BB7", "B54BB9D8", "121D4ECA", "1C0780C8", "70EC99CE"])
def test_cli_matches_golden(uuid, tmp_path):
    out = _run_cli(dict(CASES[uuid]["params"]), str(tmp_path))
    _check_golden(uuid, str(tmp_path))
    assert "Performance:" in out and "ns/gp/eq/rhs" in out
    assert os.path.exists(tmp_path / "time_data.dat") and os.path.exists(tmp_path / "simulation.inp")
    assert os.path.isdir(tmp_path / "p_all" / "p0")


def test_segments_equal_single_steps(tmp_path):
    """Whole segments replayed from the captured CUDA graph (no observer between prints) == one s_tvd_rk call per step."""
    case = dict(CASES["A0B82851"]["params"], t_step_stop=40, t_step_save=40)
    _run_cli(dict(case, t_step_print=1, run_time_info="T"), str(tmp_path / "a"))   # observed every step
    _run_cli(dict(case, t_step_print=20, run_time_info="F"), str(tmp_path / "b"))  # 10 timed steps, then segments
    a = data_io.read_serial_restart(str(tmp_path / "a"), 0, 40, 24, 24, 24, 6)[1]
    b = data_io.read_serial_restart(str(tmp_path / "b"), 0, 40, 24, 24, 24, 6)[1]
    assert np.array_equal(a, b)
    rows = open(tmp_path / "a" / "run_time.inf").read()
    assert "ICFL Max:" in rows


def test_restart_from_files(tmp_path):
    case = dict(CASES["04E1CB26"]["params"])
    _run_cli(dict(case), str(tmp_path / "a"))
    _run_cli(dict(case, t_step_stop=25, t_step_save=25), str(tmp_path / "b"))
    _run_cli(dict(case, t_step_start=25, t_step_stop=50, t_step_save=25), str(tmp_path / "b"), extra=["-t", "simulation"])
    n = 2 * int(case["num_fluids"]) + 4
    a = data_io.read_serial_restart(str(tmp_path / "a"), 0, 50, 24, 24, 24, n)[1]
    b = data_io.read_serial_restart(str(tmp_path / "b"), 0, 50, 24, 24, 24, n)[1]
    ok, amax, worst = compare(b, a, 1e-12)
    assert ok, (amax, worst)


@pytest.mark.parametrize("uuid", ["0FCCE9F1", "8C7AA13B", "CE232828"])
def test_two_ranks_match_golden(uuid, tmp_path):
    """``torchrun --nproc-per-node 2 -m mfc_b200 run``: the reference's 2-rank goldens (per-rank D/ files)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if uuid not in CASES:
        pytest.skip("fixture not present")
    _run_cli(dict(CASES[uuid]["params"]), str(tmp_path), nranks=2)
    _check_golden(uuid, str(tmp_path))


def test_cfl_adaptive_dt(tmp_path):
    """cfl_adap_dt through the CLI: dt from the device reduction (s_compute_dt); the ICFL the same step reports equals cfl_target."""
    case = dict(CASES["A0B82851"]["params"], cfl_adap_dt="T", cfl_target=0.3, t_stop=0.05, t_save=0.025, n_start=0, run_time_info="T")
    out = _run_cli(case, str(tmp_path))
    rows = [l for l in open(tmp_path / "run_time.inf").read().split("\n") if l.strip() and l.split()[0].isdigit()]
    icfl = np.array([float(r.split()[3]) for r in rows])
    dts = np.array([float(r.split()[1]) for r in rows])
    assert len(rows) > 5 and np.all(np.abs(icfl[:-1] - 0.3) < 2e-6) and abs(dts.sum() - 0.05) < 1e-5
    assert os.path.isdir(tmp_path / "p_all" / "p0" / "2") and " dt = " in out
