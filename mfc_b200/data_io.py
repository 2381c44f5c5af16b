"""On-disk formats either side of the hot path.

Restates, for the 5-equation model, what the reference reads and writes around ``s_tvd_rk``:

  * serial restart/IC data ``<case_dir>/p_all/p<rank>/<t_step>/{x,y,z}_cb.dat, q_cons_vf<i>.dat`` --
    Fortran *unformatted sequential* files: one record = 4-byte length, payload, 4-byte length
    (gfortran/nvfortran default), ``x_cb(-1:m)`` and ``q_cons_vf(i)%sf(0:m,0:n,0:p)`` column-major.
    Writer: ``src/simulation/m_data_output.fpp:294-342`` (pre_process: ``src/pre_process/m_data_output.fpp:98-120``);
    reader: ``src/simulation/m_start_up.fpp:179-266``.
  * the formatted ``<case_dir>/D/{cons,prim}.<i>.<rank:02d>.<t_step:06d>.dat`` files the regression
    packer consumes (``F40.14`` columns ``x_cb(j) [y_cb(k) [z_cb(l)]] value``, x outer / z inner, blank
    line after every innermost loop): ``src/simulation/m_data_output.fpp:386-647``.  In the ``prim`` files
    the partial densities and volume fractions are the *conservative* values (``:419-423,537-542,632-637``).
  * parallel-I/O restart data ``<case_dir>/restart_data/lustre_<t_step>.dat`` (``sys_size`` global arrays,
    x fastest, no record markers, every rank writing its sub-block) and ``lustre_{x,y,z}_cb.dat``:
    ``src/simulation/m_data_output.fpp:651-860``, ``src/simulation/m_start_up.fpp:269-506``.
  * ``run_time.inf`` (``m_data_output.fpp:66-109,248-266,1553-1572``), ``time_data.dat`` /
    ``io_time_data.dat`` (``m_start_up.fpp:674-699``).
  * ``pack_D``: the packer's rule for turning ``D/`` into golden lines
    (``toolchain/mfc/packer/pack.py:118-160``).
Host arrays are ``(N, nz, ny, nx)`` float64 (x contiguous), which *is* Fortran's ``sf(0:m,0:n,0:p)`` layout.
"""
import datetime
import os
import re
import shutil
import struct

import numpy as np


# ---- Fortran unformatted sequential records -----------------------------------------------------------
def write_record(f, arr):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    n = a.nbytes
    if n >= 2 ** 31:  # gfortran splits such records into sub-records; the reference never writes them serially
        raise ValueError("record larger than 2 GiB: use parallel_io")
    f.write(struct.pack("<i", n))
    f.write(a.tobytes())
    f.write(struct.pack("<i", n))


def read_record(f, count=None):
    head = f.read(4)
    if len(head) != 4:
        raise IOError("unexpected end of file in a Fortran record header")
    (n,) = struct.unpack("<i", head)
    if n < 0 or n % 8:
        raise IOError(f"bad Fortran record length {n}")
    data = f.read(n)
    tail = f.read(4)
    if len(data) != n or len(tail) != 4 or struct.unpack("<i", tail)[0] != n:
        raise IOError("corrupt Fortran record")
    a = np.frombuffer(data, dtype="<f8").copy()
    if count is not None and a.size != count:
        raise IOError(f"record holds {a.size} values, expected {count}")
    return a


def step_dir(case_dir, rank, t_step):
    return os.path.join(case_dir, "p_all", f"p{rank}", f"{t_step}")


def write_serial_restart(case_dir, rank, t_step, cbs, cons):
    """``p_all/p<rank>/<t_step>/``: grid + conservative variables (existing directory replaced, ``:312-315``)."""
    d = step_dir(case_dir, rank, t_step)
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d)
    for name, cb in zip("xyz", cbs):
        if cb is None:
            continue
        with open(os.path.join(d, f"{name}_cb.dat"), "wb") as f:
            write_record(f, cb)
    for i in range(cons.shape[0]):
        with open(os.path.join(d, f"q_cons_vf{i + 1}.dat"), "wb") as f:
            write_record(f, cons[i])


def read_serial_restart(case_dir, rank, t_step, m, n, p, sys_size):
    """-> ([x_cb, y_cb, z_cb], cons (N, nz, ny, nx)); missing files abort like ``s_mpi_abort`` (``m_start_up.fpp:200-262``)."""
    d = step_dir(case_dir, rank, t_step)
    if not os.path.isdir(d):
        raise IOError(f"{d} is missing. Exiting.")
    cbs = []
    for name, s in zip("xyz", (m, n, p)):
        if name != "x" and s == 0:
            cbs.append(None)
            continue
        path = os.path.join(d, f"{name}_cb.dat")
        if not os.path.exists(path):
            raise IOError(f"{path} is missing. Exiting.")
        with open(path, "rb") as f:
            cbs.append(read_record(f, s + 2))
    cons = np.empty((sys_size, p + 1, n + 1, m + 1))
    for i in range(sys_size):
        path = os.path.join(d, f"q_cons_vf{i + 1}.dat")
        if not os.path.exists(path):
            raise IOError(f"{path} is missing. Exiting.")
        with open(path, "rb") as f:
            cons[i] = read_record(f, (m + 1) * (n + 1) * (p + 1)).reshape(p + 1, n + 1, m + 1)
    return cbs, cons


# ---- formatted D/ files --------------------------------------------------------------------------------
def _write_field_text(path, cbs, a):
    """One D/ file: a is (nz, ny, nx); columns are the cell's *right* boundaries x_cb(j) = cb[j+1]."""
    nz, ny, nx = a.shape
    x = cbs[0][1:]
    with open(path, "w") as f:
        if ny == 1 and nz == 1:
            f.write("".join("%40.14f%40.14f\n" % (x[j], a[0, 0, j]) for j in range(nx)))
        elif nz == 1:
            y = cbs[1][1:]
            for j in range(nx):
                f.write("".join("%40.14f%40.14f%40.14f\n" % (x[j], y[k], a[0, k, j]) for k in range(ny)))
                f.write("\n")
        else:
            y, z = cbs[1][1:], cbs[2][1:]
            for j in range(nx):
                for k in range(ny):
                    f.write("".join("%40.14f%40.14f%40.14f%40.14f\n" % (x[j], y[k], z[l], a[l, k, j]) for l in range(nz)))
                    f.write("\n")
                f.write("\n")


def write_D(case_dir, rank, t_step, cbs, cons, prim=None, num_fluids=1, prim_vars_wrt=False):
    """``D/cons.*`` always; ``D/prim.*`` for 1-D runs and when ``prim_vars_wrt`` (``m_data_output.fpp:399-427,520-547,620-645``)."""
    d = os.path.join(case_dir, "D")
    os.makedirs(d, exist_ok=True)
    N, nz, ny, nx = cons.shape
    one_d = ny == 1 and nz == 1
    if (prim_vars_wrt or one_d) and prim is not None:
        nd = 1 + (ny > 1) + (nz > 1)
        iE = num_fluids + nd
        for i in range(N):
            adv = i > iE
            cont = i < num_fluids
            # 1-D: only the volume fractions are taken from q_cons (:419); 2-D/3-D: partial densities too (:537, :632)
            use_cons = adv or (cont and not one_d)
            src = cons[i] if use_cons else prim[i]
            _write_field_text(os.path.join(d, f"prim.{i + 1}.{rank:02d}.{t_step:06d}.dat"), cbs, src)
    for i in range(N):
        _write_field_text(os.path.join(d, f"cons.{i + 1}.{rank:02d}.{t_step:06d}.dat"), cbs, cons[i])


def pack_D(case_dir):
    """{``D/<file>``: values} -- the packer's rule: the first line tells the number of coordinate columns,
    every (ndims+1)-th number is kept (``toolchain/mfc/packer/pack.py:147-153``)."""
    out = {}
    d = os.path.join(case_dir, "D")
    for name in sorted(os.listdir(d)):
        if not name.endswith(".dat"):
            continue
        with open(os.path.join(d, name)) as f:
            content = f.read()
        nums = [float(e) for e in re.sub(r"[\n\t\s]+", " ", content).strip().split(" ")]
        first = [float(e) for e in re.sub(r"[\n\t\s]+", " ", content.split("\n", 1)[0]).strip().split(" ")]
        ndims = len(first) - 1
        out["D/" + name] = np.array(nums[ndims::ndims + 1])
    return out


# ---- parallel I/O (restart_data/) ------------------------------------------------------------------------
def _lustre(case_dir, name):
    return os.path.join(case_dir, "restart_data", "lustre_" + name)


def write_parallel_grid(case_dir, glb_cbs):
    """``restart_data/lustre_{x,y,z}_cb.dat``: the global ``s_cb(-1:s_glb)`` as raw doubles (pre_process, rank 0)."""
    os.makedirs(os.path.join(case_dir, "restart_data"), exist_ok=True)
    for name, cb in zip("xyz", glb_cbs):
        if cb is not None:
            np.ascontiguousarray(cb, dtype="<f8").tofile(_lustre(case_dir, f"{name}_cb.dat"))


def read_parallel_grid(case_dir, m_glb, n_glb, p_glb):
    cbs = []
    for name, s in zip("xyz", (m_glb, n_glb, p_glb)):
        if name != "x" and s == 0:
            cbs.append(None)
            continue
        path = _lustre(case_dir, f"{name}_cb.dat")
        if not os.path.exists(path):
            raise IOError(f"File {path} is missing. Exiting.")
        a = np.fromfile(path, dtype="<f8")
        if a.size != s + 2:
            raise IOError(f"{path}: {a.size} values, expected {s + 2}")
        cbs.append(a)
    return cbs


def create_parallel_restart(case_dir, t_step, glb_mnp, sys_size):
    """Create (or reset to the right size) ``restart_data/lustre_<t_step>.dat``.  ONE rank calls this, followed by a
    barrier, before the ranks write their blocks (MPI_FILE_OPEN with MPI_MODE_CREATE is collective in the reference)."""
    m, n, p = glb_mnp
    path = _lustre(case_dir, f"{t_step}.dat")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "ab") as f:
        f.truncate(sys_size * (p + 1) * (n + 1) * (m + 1) * 8)


def _parallel_map(case_dir, t_step, glb_shape, sys_size, mode):
    nz, ny, nx = glb_shape
    path = _lustre(case_dir, f"{t_step}.dat")
    if not os.path.exists(path):
        raise IOError(f"File {path} is missing. Exiting.")
    return np.memmap(path, dtype="<f8", mode="r" if mode == "r" else "r+", shape=(sys_size, nz, ny, nx))


def write_parallel_restart(case_dir, t_step, glb_mnp, start, cons):
    """This rank's block of ``restart_data/lustre_<t_step>.dat`` (the MPI-IO subarray view, ``m_data_output.fpp:795-830``).
    ``start`` = (i0, j0, k0) global index of the block's first cell.  Ranks write disjoint regions of one file, which
    ``create_parallel_restart`` must have created first."""
    m, n, p = glb_mnp
    N, nz, ny, nx = cons.shape
    mm = _parallel_map(case_dir, t_step, (p + 1, n + 1, m + 1), N, "w")
    i0, j0, k0 = start
    mm[:, k0:k0 + nz, j0:j0 + ny, i0:i0 + nx] = cons
    mm.flush()
    del mm


def read_parallel_restart(case_dir, t_step, glb_mnp, start, loc_mnp, sys_size):
    m, n, p = glb_mnp
    lm, ln, lp = loc_mnp
    mm = _parallel_map(case_dir, t_step, (p + 1, n + 1, m + 1), sys_size, "r")
    i0, j0, k0 = start
    out = np.array(mm[:, k0:k0 + lp + 1, j0:j0 + ln + 1, i0:i0 + lm + 1], dtype=np.float64)
    del mm
    return out


# ---- run_time.inf, time_data.dat -------------------------------------------------------------------------
class RunTimeInfo:
    """``run_time.inf`` of rank 0 (``s_open/_write/_close_run_time_information``)."""

    def __init__(self, case_dir):
        self.icfl_max = 0.0
        self.f = open(os.path.join(case_dir, "run_time.inf"), "w")
        w = self.f.write
        w("Description: Stability information at each time-step of the simulation. This\n")
        for s in ("data is composed of the inviscid Courant-Friedrichs-Lewy (ICFL)", "number, the viscous CFL (VCFL) number, the capillary CFL (CCFL)",
                  "number and the cell Reynolds (Rc) number. Please note that only", "those stability conditions pertinent to the physics included in",
                  "the current computation are displayed."):
            w(" " * 13 + s + "\n")
        w("Date: " + datetime.date.today().strftime("%m/%d/%y") + "\n\n\n")
        w(" " * 13 + "Time-step" + " " * 13 + "%10s" % "dt" + " " * 13 + "%10s" % "Time" + " " * 13 + "%10s" % "ICFL Max" + "\n")

    def write(self, t_step, dt, mytime, icfl):
        """One row; raises like ``s_mpi_abort`` on a NaN or > 1 ICFL (``m_data_output.fpp:268-273``)."""
        self.icfl_max = max(self.icfl_max, icfl) if icfl == icfl else self.icfl_max
        self.f.write(" " * 13 + "%9d" % t_step + " " * 13 + "%10.6f" % dt + " " * 13 + "%10.6f" % mytime + " " * 13 + "%10.6f" % icfl + "\n")
        if icfl != icfl:
            self.f.flush()
            raise FloatingPointError("ICFL is NaN. Exiting.")
        if icfl > 1.0:
            self.f.flush()
            raise FloatingPointError(f"icfl {icfl}: ICFL is greater than 1.0. Exiting.")

    def close(self, run_time_s):
        self.f.write("    \n\n")
        self.f.write("ICFL Max: %9.6f\n" % self.icfl_max)
        self.f.write("\nRun-time: %ds\n    \n" % int(round(run_time_s)))
        self.f.close()


def append_time_data(case_dir, num_procs, time_final, grind_time, io_time_final):
    """``time_data.dat`` / ``io_time_data.dat`` rows (``m_start_up.fpp:678-699``); created with their header if new."""
    p = os.path.join(case_dir, "time_data.dat")
    new = not os.path.exists(p)
    with open(p, "a") as f:
        if new:
            f.write("%10s%15s%15s\n" % ("Ranks", "s/step", "ns/gp/eq/rhs"))
        f.write("%10d%15.8f%15.8f\n" % (num_procs, time_final, grind_time))
    p = os.path.join(case_dir, "io_time_data.dat")
    new = not os.path.exists(p)
    with open(p, "a") as f:
        if new:
            f.write("%10s%15s\n" % ("Ranks", "s/step"))
        f.write("%10d%15.8f\n" % (num_procs, io_time_final))
