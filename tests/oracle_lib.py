"""ctypes binding of the CPU oracle (oracle/libmfc5_oracle.so) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from mfc_b200.params import Mfc5Params

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORC_DIR = os.path.join(_ROOT, "oracle")
_dp = C.POINTER(C.c_double)
_libs = {}


def load(omp=False):
    name = "libmfc5_oracle_omp.so" if omp else "libmfc5_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_ORC_DIR, name)
    src = os.path.join(_ORC_DIR, "mfc5_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _ORC_DIR], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(Mfc5Params)]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_sys_size.argtypes = [C.c_void_p]
    L.orc_buff_size.argtypes = [C.c_void_p]
    L.orc_set_grid.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    L.orc_finalize_setup.argtypes = [C.c_void_p]
    L.orc_set_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    for f in ("orc_set_cons", "orc_get_cons", "orc_get_prim", "orc_get_rhs"):
        getattr(L, f).argtypes = [C.c_void_p, _dp]
    L.orc_box_size.argtypes = [C.c_void_p]
    L.orc_box_size.restype = C.c_long
    L.orc_get_box.argtypes = [C.c_void_p, C.c_int, _dp]
    L.orc_set_cons_box.argtypes = [C.c_void_p, _dp]
    L.orc_cons2prim.argtypes = [C.c_void_p]
    L.orc_bc_fill.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_halo_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_halo_count.restype = C.c_int
    L.orc_halo_pack.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.orc_halo_unpack.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.orc_sweeps.argtypes = [C.c_void_p]
    L.orc_rk_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.orc_compute_rhs.argtypes = [C.c_void_p]
    L.orc_tvd_rk_step.argtypes = [C.c_void_p, C.c_double]
    L.orc_set_time.argtypes = [C.c_void_p, C.c_double]
    L.orc_icfl_max.argtypes = [C.c_void_p, C.c_double]
    L.orc_icfl_max.restype = C.c_double
    L.orc_num_threads.restype = C.c_int
    _libs[name] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class Oracle:
    """One rank's state in the CPU oracle.  Arrays are (N, nz, ny, nx) float64, interior only."""

    def __init__(self, P, cbs, ghost_widths=None, omp=False):
        self.L = load(omp)
        self.P = P
        self.h = self.L.orc_create(C.byref(P))
        self.N = self.L.orc_sys_size(self.h)
        self.b = self.L.orc_buff_size(self.h)
        self.shape = (self.N, P.p + 1, P.n + 1, P.m + 1)
        for d in range(P.num_dims):
            cb = np.ascontiguousarray(cbs[d], dtype=np.float64)
            gb = ge = None
            if ghost_widths is not None and ghost_widths[d] is not None:
                gb, ge = ghost_widths[d]
                gb = None if gb is None else np.ascontiguousarray(gb, dtype=np.float64)
                ge = None if ge is None else np.ascontiguousarray(ge, dtype=np.float64)
            self.L.orc_set_grid(self.h, d, _ptr(cb), _ptr(gb), _ptr(ge))
        self.L.orc_finalize_setup(self.h)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_dirichlet(self, d, side, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self.L.orc_set_dirichlet(self.h, d, side, _ptr(buf))

    def set_cons(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.shape, (q.shape, self.shape)
        self.L.orc_set_cons(self.h, _ptr(q))

    def _get(self, fn):
        q = np.empty(self.shape)
        fn(self.h, _ptr(q))
        return q

    def get_cons(self):
        return self._get(self.L.orc_get_cons)

    def get_prim(self):
        return self._get(self.L.orc_get_prim)

    def get_rhs(self):
        return self._get(self.L.orc_get_rhs)

    @property
    def box_shape(self):
        """(N, p+1+2bz, n+1+2by, m+1+2b): the buffered box sf(-b:m+b, -b:n+b, -b:p+b), collapsed dims 0:0."""
        P, b = self.P, self.b
        return (self.N, P.p + 1 + (2 * b if P.p > 0 else 0), P.n + 1 + (2 * b if P.n > 0 else 0), P.m + 1 + 2 * b)

    def get_box(self, which):
        """which = 'cons' | 'prim' | 'rhs': the buffered box as the reference holds it (ghost cells as last filled)."""
        q = np.empty(self.box_shape)
        assert q.size == self.N * self.L.orc_box_size(self.h)
        self.L.orc_get_box(self.h, {"cons": 0, "prim": 1, "rhs": 2}[which], _ptr(q))
        return q

    def set_cons_box(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.box_shape
        self.L.orc_set_cons_box(self.h, _ptr(q))

    def compute_rhs(self):
        self.L.orc_compute_rhs(self.h)

    def step(self, dt):
        self.L.orc_tvd_rk_step(self.h, dt)

    def set_time(self, mytime):
        self.L.orc_set_time(self.h, mytime)

    def icfl_max(self, dt):
        return self.L.orc_icfl_max(self.h, dt)

    # phase API for multi-rank tests
    def cons2prim(self):
        self.L.orc_cons2prim(self.h)

    def bc_fill(self, d, side):
        self.L.orc_bc_fill(self.h, d, side)

    def halo_pack(self, d, side):
        buf = np.empty(self.L.orc_halo_count(self.h, d))
        self.L.orc_halo_pack(self.h, d, side, _ptr(buf))
        return buf

    def halo_unpack(self, d, side, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self.L.orc_halo_unpack(self.h, d, side, _ptr(buf))

    def sweeps(self):
        self.L.orc_sweeps(self.h)

    def rk_update(self, stage, dt):
        self.L.orc_rk_update(self.h, stage, dt)
