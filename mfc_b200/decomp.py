"""Cartesian block decomposition of the computational domain over ranks (one rank = one GPU).

Restates ``s_mpi_decompose_computational_domain`` (reference ``src/common/m_mpi_common.fpp:1319-1678``):
the processor-grid factorisation that minimises ``|(m+1)/px - (n+1)/py| + |(n+1)/py - (p+1)/pz|`` (real
division, ``>=`` so the LAST admissible minimum in loop order wins, ``:1426-1462``; 2-D ``:1521-1547``),
the remainder cells going to the low ranks (``:1482-1493,1566-1577,1619-1630``), ``MPI_CART_RANK``
row-major rank numbering, and the neighbour rank stored in ``bc_<dir>%beg/end`` as a non-negative value
(``:1496-1509,1580-1593,1635-1648``; periodic directions split over more than one rank become
neighbour links).  Local grids follow ``s_apply_decomposition_policies`` (``:1681-1740``, serial I/O:
per-rank domain bounds, uniform spacing) or a slice of the global grid (``parallel_io = T``,
``s_apply_grid_from_global_dim`` ``:1800``); the ``buff_size`` neighbour cell widths that
``s_mpi_sendrecv_grid_variable_buffer`` (``:1745-1796``) would deliver are taken from the neighbour's
grid directly, since the host driver knows the global grid.
"""
import numpy as np

from . import pre_process
from .params import UnsupportedCase, _logical

NUM_STCLS_MIN = 5  # src/common/m_constants.fpp


def factorize(nranks, m, n, p, recon_order):
    """(px, py, pz) exactly as the reference chooses it."""
    need = NUM_STCLS_MIN * recon_order
    if n == 0:
        return (nranks, 1, 1)
    if p == 0:
        px, py, ok = 1, nranks, False
        fct_min = 10.0 * abs((m + 1) / float(px) - (n + 1) / float(py))
        for i in range(1, nranks + 1):
            if nranks % i == 0 and (m + 1) / float(i) >= need:
                tx, ty = float(i), float(nranks // i)
                f = abs((m + 1) / tx - (n + 1) / ty)
                if fct_min >= f and (n + 1) / ty >= need:
                    px, py, fct_min, ok = i, nranks // i, f, True
        if not ok:
            raise UnsupportedCase("Unsupported combination of values of num_procs, m, n and weno_order")
        return (px, py, 1)
    px, py, pz, ok = 1, 1, nranks, False
    fct_min = 10.0 * abs((m + 1) / 1.0 - (n + 1) / 1.0) + 10.0 * abs((n + 1) / 1.0 - (p + 1) / float(nranks))
    for i in range(1, nranks + 1):
        if nranks % i == 0 and (m + 1) / float(i) >= need:
            for j in range(1, nranks // i + 1):
                if (nranks // i) % j == 0 and (n + 1) / float(j) >= need:
                    tx, ty, tz = float(i), float(j), float(nranks // (i * j))
                    f = abs((m + 1) / tx - (n + 1) / ty) + abs((n + 1) / ty - (p + 1) / tz)
                    if fct_min >= f and (p + 1) / tz >= need:
                        px, py, pz, fct_min, ok = i, j, nranks // (i * j), f, True
    if not ok:
        raise UnsupportedCase("Unsupported combination of values of num_procs, m, n, p and weno_order")
    return (px, py, pz)


def cart_coords(rank, dims, nd):
    """MPI_CART_COORDS for a row-major Cartesian communicator over the first ``nd`` entries of dims."""
    c = [0, 0, 0]
    r = rank
    for d in range(nd - 1, -1, -1):
        c[d] = r % dims[d]
        r //= dims[d]
    return c


def cart_rank(coords, dims, nd):
    r = 0
    for d in range(nd):
        r = r * dims[d] + (coords[d] % dims[d])
    return r


def split(s_glb, nprocs, coord):
    """(start index, local cells-1) of block ``coord`` along a direction with ``s_glb + 1`` cells."""
    rem = (s_glb + 1) % nprocs
    s = (s_glb + 1) // nprocs - 1
    if coord < rem:
        return (s + 2) * coord, s + 1
    return (s + 1) * coord + rem, s


def _recon_order(case):
    """weno_order, or muscl_order for recon_type = 2 (m_checker.fpp:52-69 uses the reconstruction order)."""
    rt = case.get("recon_type", 1)
    if rt in (2, "muscl"):
        mo = case.get("muscl_order", 1)
        return {"first_order": 1, "second_order": 2}.get(mo, mo)
    return int(case["weno_order"])


class Block:
    """Everything one rank needs: local extents, neighbour ranks, local grid and neighbour cell widths."""

    def __init__(self, case, rank, nranks, dims=None):
        m, n, p = int(case["m"]), int(case.get("n", 0)), int(case.get("p", 0))
        self.glb = (m, n, p)
        self.nd = 1 + (n > 0) + (p > 0)
        self.rank, self.nranks = rank, nranks
        self.dims = tuple(dims) if dims is not None else factorize(nranks, m, n, p, _recon_order(case))
        assert self.dims[0] * self.dims[1] * self.dims[2] == nranks
        self.coords = cart_coords(rank, self.dims, self.nd)
        self.start, self.s = [0, 0, 0], [m, n, p]
        for d in range(self.nd):
            self.start[d], self.s[d] = split(self.glb[d], self.dims[d], self.coords[d])
        # boundary conditions: physical codes kept on the domain boundary, neighbour ranks elsewhere
        self.bc = [[0, 0], [0, 0], [0, 0]]
        for d, x in enumerate("xyz"):
            if d >= self.nd:
                continue
            beg, end = int(case[f"bc_{x}%beg"]), int(case[f"bc_{x}%end"])
            c = list(self.coords)
            if self.coords[d] > 0 or (beg == -1 and self.dims[d] > 1):
                c[d] = self.coords[d] - 1
                beg = cart_rank(c, self.dims, self.nd)
            c = list(self.coords)
            if self.coords[d] < self.dims[d] - 1 or (end == -1 and self.dims[d] > 1):
                c[d] = self.coords[d] + 1
                end = cart_rank(c, self.dims, self.nd)
            self.bc[d] = [beg, end]

    def local_case(self, case):
        """The case dictionary of this rank: local m, n, p and bc (what the reference holds after
        s_mpi_decompose_computational_domain)."""
        c = dict(case)
        c["m"], c["n"], c["p"] = self.s
        for d, x in enumerate("xyz"):
            if d < self.nd:
                c[f"bc_{x}%beg"], c[f"bc_{x}%end"] = self.bc[d]
        return c

    def grid(self, case, b):
        """(cbs, ccs, ghost_widths) of this rank.  ghost_widths[d] = (beg, end) arrays of ``b`` neighbour
        cell widths ordered by increasing index (None on physical boundaries)."""
        cbs, ccs, gws = [], [], []
        glb_cbs, glb_ccs = pre_process.generate_grid(case)
        par_io = _logical(case.get("parallel_io", "F"))
        for d, x in enumerate("xyz"):
            if d >= self.nd:
                cbs.append(None)
                ccs.append(np.zeros(1))
                gws.append(None)
                continue
            st, s = self.start[d], self.s[d]
            stretched = _logical(case.get(f"stretch_{x}", "F"))
            gcb = glb_cbs[d]
            if par_io or stretched or self.dims[d] == 1:
                cb = gcb[st:st + s + 2].copy()
                cc = glb_ccs[d][st:st + s + 1].copy()
            else:
                # serial I/O: pre_process generates a uniform grid on the rank's own domain bounds
                dbeg, dend = float(case[f"{x}_domain%beg"]), float(case[f"{x}_domain%end"])
                spacing = (dend - dbeg) / float(self.glb[d] + 1)
                nproc, coord = self.dims[d], self.coords[d]
                rem = (self.glb[d] + 1) % nproc
                if coord < rem:
                    lbeg = dbeg + spacing * float((s + 1) * coord)
                    lend = dend - spacing * float((s + 1) * (nproc - coord - 1) - (nproc - rem))
                else:
                    lbeg = dbeg + spacing * float((s + 1) * coord + rem)
                    lend = dend - spacing * float((s + 1) * (nproc - coord - 1))
                cb = pre_process.generate_grid_1d(case, x, s, lbeg, lend)
                dx = (lend - lbeg) / float(s + 1)
                cc = np.array([lbeg + 5.0e-1 * dx * float(2 * i + 1) for i in range(s + 1)])
            gds = gcb[1:] - gcb[:-1]  # global cell widths, index 0..s_glb
            nglb = self.glb[d] + 1
            gb = ge = None
            if self.bc[d][0] >= 0:
                gb = np.array([gds[(st - b + i) % nglb] for i in range(b)])
            if self.bc[d][1] >= 0:
                ge = np.array([gds[(st + s + 1 + i) % nglb] for i in range(b)])
            cbs.append(cb)
            ccs.append(cc)
            gws.append((gb, ge))
        return cbs, ccs, gws

    def slab(self, arr):
        """Slice a global (N, nz, ny, nx) array to this rank's block."""
        sl = [slice(None)]
        for d in (2, 1, 0):
            sl.append(slice(self.start[d], self.start[d] + self.s[d] + 1) if d < self.nd else slice(None))
        return np.ascontiguousarray(arr[tuple(sl)])
