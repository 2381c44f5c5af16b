"""ctypes binding of the C-ABI CUDA library ``libmfc5.so`` (``include/mfc5.h``).

``Mfc5`` is the host-side mirror of the reference's module state for the hot path: it owns one
handle (one rank / one GPU), uploads ``q_cons_vf`` once and then advances it with ``s_tvd_rk``
(reference ``src/simulation/m_time_steppers.fpp:438``) entirely on the device.  There is no CPU
fallback: if the CUDA library cannot be loaded, or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from .params import Mfc5Params

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_libs = {}


class Mfc5Error(RuntimeError):
    pass


def load(strict=False):
    """Load libmfc5.so (``strict``: the -fmad=false build used by the parity tests)."""
    name = "libmfc5_strict.so" if strict else "libmfc5.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not strict and os.environ.get("MFC5_LIB"):  # kernel-variant experiments (tools/sweep_bench.py)
        path = os.environ["MFC5_LIB"]
    if not os.path.exists(path):
        raise Mfc5Error(f"{path} is missing: build it with `make -C mfc_b200/csrc` (or __graft_entry__.build()); "
                        "there is no CPU fallback")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.mfc5_create.argtypes = [C.POINTER(Mfc5Params), C.POINTER(vp)]
    L.mfc5_destroy.argtypes = [vp]
    L.mfc5_set_grid.argtypes = [vp, C.c_int, _dp, _dp, _dp]
    L.mfc5_finalize_setup.argtypes = [vp]
    L.mfc5_set_dirichlet.argtypes = [vp, C.c_int, C.c_int, _dp]
    for f in ("mfc5_upload_cons", "mfc5_download_cons", "mfc5_download_prim", "mfc5_download_rhs"):
        getattr(L, f).argtypes = [vp, C.POINTER(_dp), C.c_int]
    L.mfc5_compute_rhs.argtypes = [vp]
    L.mfc5_tvd_rk.argtypes = [vp, C.c_double, _dp]
    L.mfc5_tvd_rk_n.argtypes = [vp, C.c_double, C.c_int]
    L.mfc5_sync.argtypes = [vp]
    L.mfc5_graph_state.argtypes = [vp]
    L.mfc5_set_time.argtypes = [vp, C.c_double]
    L.mfc5_reduce.argtypes = [vp, C.c_int, C.c_double, _dp]
    L.mfc5_nccl_unique_id.argtypes = [C.c_char_p]
    L.mfc5_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.mfc5_sys_size.argtypes = [vp]
    L.mfc5_buff_size.argtypes = [vp]
    L.mfc5_kernel_launches.argtypes = [vp]
    L.mfc5_kernel_launches.restype = C.c_longlong
    L.mfc5_last_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong)]
    L.mfc5_last_kernel_ms.restype = C.c_double
    L.mfc5_compute_stream.argtypes = [vp]
    L.mfc5_compute_stream.restype = vp
    L.mfc5_last_error.restype = C.c_char_p
    L.mfc5_version.restype = C.c_char_p
    _libs[name] = L
    return L


EXPORTED_SYMBOLS = [
    "mfc5_create", "mfc5_destroy", "mfc5_set_grid", "mfc5_finalize_setup", "mfc5_set_dirichlet", "mfc5_upload_cons",
    "mfc5_download_cons", "mfc5_download_prim", "mfc5_download_rhs", "mfc5_compute_rhs", "mfc5_tvd_rk", "mfc5_tvd_rk_n",
    "mfc5_sync", "mfc5_graph_state", "mfc5_set_time", "mfc5_reduce", "mfc5_nccl_unique_id", "mfc5_comm_init", "mfc5_sys_size", "mfc5_buff_size",
    "mfc5_kernel_launches", "mfc5_last_kernel_ms", "mfc5_compute_stream", "mfc5_last_error", "mfc5_version",
]


KERNEL_CLASSES = ["cons2prim", "bc_fill", "halo", "cbc", "sweep_x", "sweep_y", "sweep_z"]


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class Mfc5:
    """One rank's device-resident state.  Host arrays are (N, nz, ny, nx) float64, interior only."""

    def __init__(self, P, cbs, ghost_widths=None, dirichlet=None, strict=False):
        self.L = load(strict)
        self.P = P
        self.h = C.c_void_p()
        self._ck(self.L.mfc5_create(C.byref(P), C.byref(self.h)))
        self.N = self.L.mfc5_sys_size(self.h)
        self.b = self.L.mfc5_buff_size(self.h)
        self.shape = (self.N, P.p + 1, P.n + 1, P.m + 1)
        for d in range(P.num_dims):
            cb = np.ascontiguousarray(cbs[d], dtype=np.float64)
            gb = ge = None
            if ghost_widths is not None and ghost_widths[d] is not None:
                gb, ge = ghost_widths[d]
                gb = None if gb is None else np.ascontiguousarray(gb, dtype=np.float64)
                ge = None if ge is None else np.ascontiguousarray(ge, dtype=np.float64)
            self._ck(self.L.mfc5_set_grid(self.h, d, _ptr(cb), _ptr(gb), _ptr(ge)))
        self._ck(self.L.mfc5_finalize_setup(self.h))
        for (d, side), buf in (dirichlet or {}).items():
            buf = np.ascontiguousarray(buf, dtype=np.float64)
            self._ck(self.L.mfc5_set_dirichlet(self.h, d, side, _ptr(buf)))
        self.stage_time_min = 0.0

    def _ck(self, rc):
        if rc != 0:
            raise Mfc5Error(self.L.mfc5_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mfc5_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _field_ptrs(self, q):
        arr = (_dp * self.N)()
        for v in range(self.N):
            arr[v] = q[v].ctypes.data_as(_dp)
        return arr

    def set_cons(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.shape, (q.shape, self.shape)
        self._ck(self.L.mfc5_upload_cons(self.h, self._field_ptrs(q), 0))

    def _get(self, fn):
        q = np.empty(self.shape)
        self._ck(fn(self.h, self._field_ptrs(q), 0))
        return q

    def get_cons(self):
        return self._get(self.L.mfc5_download_cons)

    def get_prim(self):
        """Primitive variables of the CURRENT state (the library converts first if its buffer is stale, e.g. after a step)."""
        return self._get(self.L.mfc5_download_prim)

    def get_rhs(self):
        return self._get(self.L.mfc5_download_rhs)

    # ---- the layout the Fortran host passes: q_cons_vf(i)%sf(-b:m+b, -b:n+b, -b:p+b) (``ghosted = 1``) ----
    @property
    def box_shape(self):
        """(N, p+1+2bz, n+1+2by, m+1+2b): the buffered box of every field (collapsed dimensions have no ghosts)."""
        P, b = self.P, self.b
        return (self.N, P.p + 1 + (2 * b if P.p > 0 else 0), P.n + 1 + (2 * b if P.n > 0 else 0), P.m + 1 + 2 * b)

    def set_cons_box(self, q):
        """Upload from buffered arrays (only the interior is read, like GPU_UPDATE of q_cons_ts(1))."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == self.box_shape, (q.shape, self.box_shape)
        self._ck(self.L.mfc5_upload_cons(self.h, self._field_ptrs(q), 1))

    def get_box(self, which, out=None):
        """Download into buffered arrays: 'cons' / 'rhs' (interior + whatever the device ghosts hold) or 'prim'
        (ghost cells filled in the reference's x -> y -> z order, edges and corners included, m_rhs.fpp:884-898)."""
        fn = {"cons": self.L.mfc5_download_cons, "prim": self.L.mfc5_download_prim, "rhs": self.L.mfc5_download_rhs}[which]
        q = np.zeros(self.box_shape) if out is None else out
        self._ck(fn(self.h, self._field_ptrs(q), 1))
        return q

    def compute_rhs(self):
        """``s_compute_rhs`` (reference m_rhs.fpp:574): RHS of the current state, no update."""
        self._ck(self.L.mfc5_compute_rhs(self.h))

    def step(self, dt):
        """``s_tvd_rk`` (reference m_time_steppers.fpp:438): one step, grind timer updated."""
        t = C.c_double(self.stage_time_min)
        self._ck(self.L.mfc5_tvd_rk(self.h, dt, C.byref(t)))
        self.stage_time_min = t.value

    def steps(self, dt, n):
        self._ck(self.L.mfc5_tvd_rk_n(self.h, dt, n))

    def sync(self):
        self._ck(self.L.mfc5_sync(self.h))

    @property
    def graph_state(self):
        """1: ``steps`` replays a CUDA graph; 0: eager by request; -1: graph capture failed, eager fallback."""
        return int(self.L.mfc5_graph_state(self.h))

    def set_time(self, mytime):
        """The reference's ``mytime`` (only the body-force acceleration depends on it); steps advance it by dt."""
        self._ck(self.L.mfc5_set_time(self.h, mytime))

    def reduce(self, what, dt=0.0):
        out = C.c_double()
        self._ck(self.L.mfc5_reduce(self.h, what, dt, C.byref(out)))
        return out.value

    def icfl_max(self, dt):
        return self.reduce(0, dt)

    @property
    def kernel_launches(self):
        return int(self.L.mfc5_kernel_launches(self.h))

    def last_kernel_ms(self):
        """{kernel class: (device ms, launches)} of the last ``step`` (CUDA events, compute stream)."""
        out = {}
        for k, name in enumerate(KERNEL_CLASSES):
            n = C.c_longlong()
            ms = self.L.mfc5_last_kernel_ms(self.h, k, C.byref(n))
            out[name] = (ms, n.value)
        return out

    @property
    def compute_stream(self):
        return self.L.mfc5_compute_stream(self.h)

    def comm_init(self, unique_id, rank, nranks):
        self._ck(self.L.mfc5_comm_init(self.h, unique_id, rank, nranks))


def nccl_unique_id(strict=False):
    L = load(strict)
    buf = C.create_string_buffer(128)
    if L.mfc5_nccl_unique_id(buf) != 0:
        raise Mfc5Error(L.mfc5_last_error().decode())
    return buf.raw
