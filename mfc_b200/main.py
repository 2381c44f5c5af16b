"""Host driver for a stock reference case: ``pre_process`` + ``simulation`` around the CUDA hot path.

    python -m mfc_b200 run examples/1D_sodshocktube/case.py [--case-dir DIR] [-t pre_process simulation] [-- case args]
    python -m torch.distributed.run --nproc-per-node N -m mfc_b200 run case.py        # one rank per GPU (the reference: mpirun -n N)

Restates ``p_main`` (reference ``src/simulation/p_main.fpp:45-91``), ``s_perform_time_step``
(``src/simulation/m_start_up.fpp:566-642``), ``s_save_data`` (``:705-799``), ``s_save_performance_metrics``
(``:645-701``) and the ``pre_process`` main program (``src/pre_process/p_main.f90``): the same
``t_step_start / t_step_stop / t_step_save / t_step_print`` bookkeeping, the same last-step ``dt`` clipping, the
``[ xx%]  Time step ...`` progress lines, the ``Performance: ... ns/gp/eq/rhs`` line, ``time_data.dat``,
``run_time.inf`` and the data files of ``data_io``.  The time step itself is ``Mfc5.step`` = ``s_tvd_rk`` on the GPU;
there is no CPU fallback (``make_state`` exists so the CPU test-suite can drive this loop with the oracle).
"""
import argparse
import math
import os
import sys
import time

import numpy as np

from . import case_io, data_io, decomp, pre_process
from .params import _logical, params_from_case

N_WARMUP = 2      # src/simulation/m_time_steppers.fpp:450
N_TIMED = 8       # steps after the warm-up that are timed stage by stage (then whole segments are enqueued)


class World:
    """Rank bookkeeping.  Under torchrun: one process per GPU, gloo group for the few host-side collectives."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", 0))
        self.size = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.dist = None
        if self.size > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("gloo")
            self.dist = dist

    def bcast(self, obj):
        if self.dist is None:
            return obj
        box = [obj if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def max(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def _gpu_state(world):
    def make(P, cbs, gws, dirichlet):
        from . import lib
        P.device = world.local_rank
        g = lib.Mfc5(P, cbs, ghost_widths=gws, dirichlet=dirichlet)
        if world.size > 1:
            uid = world.bcast(lib.nccl_unique_id() if world.rank == 0 else None)
            g.comm_init(uid, world.rank, world.size)
        return g
    return make


def run_pre_process(case, case_dir, world=None, log=print):
    """Grid + patch initial condition -> ``p_all/p<rank>/0`` (or ``restart_data/``) and ``D/*.000000.dat``."""
    world = world or World()
    blk = decomp.Block(case, world.rank, world.size)
    lcase = blk.local_case(case)
    P = params_from_case(lcase)
    par_io = _logical(case.get("parallel_io", "F"))
    cbs, ccs, _ = blk.grid(case, P.buff_size)
    prim, cons = pre_process.initial_condition(lcase, ccs, ds_min=pre_process.grid_minima(case))
    t0 = int(case.get("t_step_start", 0))
    if _logical(case.get("cfl_adap_dt", "F")) or _logical(case.get("cfl_const_dt", "F")):
        t0 = int(case.get("n_start", 0))   # src/pre_process/m_data_output.fpp:185
    if par_io:
        if world.rank == 0:
            glb_cbs, _ = pre_process.generate_grid(case)
            data_io.write_parallel_grid(case_dir, glb_cbs)
            data_io.create_parallel_restart(case_dir, t0, blk.glb, P.sys_size)
        world.barrier()
        data_io.write_parallel_restart(case_dir, t0, blk.glb, blk.start, cons)
    else:
        data_io.write_serial_restart(case_dir, world.rank, t0, cbs, cons)
        data_io.write_D(case_dir, world.rank, t0, cbs, cons, prim, P.num_fluids, prim_vars_wrt=False)
    world.barrier()
    if world.rank == 0:
        log(f" pre_process: {blk.glb[0] + 1}x{blk.glb[1] + 1}x{blk.glb[2] + 1} cells, {world.size} rank(s), processor grid {blk.dims}")
    return prim, cons


def _progress(t_step, t0, t1, wall_avg, wall):
    # m_start_up.fpp:611-621
    pct = int(math.ceil(100.0 * (float(t_step - t0) / (t1 - t0 + 1))))
    eta = wall_avg * float(t1 - t_step)
    return (" [%3d%%]  Time step %8d of %d @ t_step = %8d Time Avg = %12.6E Time/step= %12.6E ETA (HH:MM:SS) = %d:%02d:%02d"
            % (pct, t_step - t0 + 1, t1 - t0 + 1, t_step, wall_avg, wall, int(eta) // 3600, (int(eta) % 3600) // 60, int(eta) % 60))


def run_simulation(case, case_dir, world=None, make_state=None, log=print):
    """The simulation executable: read the start data, march ``t_step_start -> t_step_stop``, save every
    ``t_step_save`` steps.  Returns a dict with the grind time and the final conservative state of this rank."""
    world = world or World()
    make_state = make_state or _gpu_state(world)
    blk = decomp.Block(case, world.rank, world.size)
    lcase = blk.local_case(case)
    P = params_from_case(lcase)
    par_io = _logical(case.get("parallel_io", "F"))
    t0, t1 = int(case.get("t_step_start", 0)), int(case.get("t_step_stop", 0))
    t_save = int(case.get("t_step_save", max(t1 - t0, 1)))
    t_print = int(case.get("t_step_print", 1))
    run_time_info = _logical(case.get("run_time_info", "F"))
    prim_wrt = _logical(case.get("prim_vars_wrt", "F"))
    dt = float(case["dt"])
    # CFL-based time stepping (cfl_adap_dt / cfl_const_dt): dt from s_compute_dt (m_time_steppers.fpp:639-710), the run is
    # bounded by t_stop and saved every t_save of simulated time; the start data sit in directory n_start (p_main.fpp:48-50)
    cfl_adap, cfl_const = _logical(case.get("cfl_adap_dt", "F")), _logical(case.get("cfl_const_dt", "F"))
    cfl_dt = cfl_adap or cfl_const
    if cfl_dt:
        t_stop_time, t_save_time = float(case["t_stop"]), float(case["t_save"])
        n_start, cfl_target = int(case.get("n_start", 0)), float(case["cfl_target"])
        t0 = n_start

    # ---- start data (s_read_data_files): grid + q_cons of t_step_start
    if par_io:
        glb_cbs = data_io.read_parallel_grid(case_dir, *blk.glb)
        cbs = [None if c is None else c[blk.start[d]:blk.start[d] + blk.s[d] + 2].copy() for d, c in enumerate(glb_cbs)]
        cons = data_io.read_parallel_restart(case_dir, t0, blk.glb, blk.start, blk.s, P.sys_size)
        gws = _ghost_widths(blk, glb_cbs, P.buff_size)
    else:
        cbs, cons = data_io.read_serial_restart(case_dir, world.rank, t0, P.m, P.n, P.p, P.sys_size)
        _, _, gws = blk.grid(case, P.buff_size)
    dirichlet = None
    if any(P.bc[d][s] == -17 for d in range(P.num_dims) for s in range(2)):
        # BC -17 buffers hold the initial primitive field on the face (m_boundary_io.fpp:271-330): regenerate it
        _, ccs, _ = blk.grid(case, P.buff_size)
        prim0, _ = pre_process.initial_condition(lcase, ccs, ds_min=pre_process.grid_minima(case))
        dirichlet = pre_process.dirichlet_buffers(P, prim0)
    state = make_state(P, cbs, gws, dirichlet)
    state.set_cons(cons)

    rti = data_io.RunTimeInfo(case_dir) if (run_time_info and world.rank == 0) else None
    mytime = 0.0 if t0 == 0 else t0 * dt           # p_main.fpp:52-58
    finaltime = t1 * dt
    t_step = t0
    time_avg = 0.0                                  # min stage time after N_WARMUP steps (m_time_steppers.fpp:571-577)
    wall_avg = wall = 0.0
    io_time, n_saves = 0.0, 0
    run_start = time.perf_counter()
    timed_api = hasattr(state, "stage_time_min")

    def save(ts):
        nonlocal io_time, n_saves
        t = time.perf_counter()
        q = state.get_cons()
        if np.isnan(q).any():                      # s_save_data, m_start_up.fpp:741-751
            raise FloatingPointError("NaN(s) in timestep output.")
        if par_io:
            if world.rank == 0:
                data_io.create_parallel_restart(case_dir, ts, blk.glb, P.sys_size)
            world.barrier()
            data_io.write_parallel_restart(case_dir, ts, blk.glb, blk.start, q)
            world.barrier()
        else:
            prim = None
            if prim_wrt or P.num_dims == 1:
                prim = _current_prim(state)
            data_io.write_serial_restart(case_dir, world.rank, ts, cbs, q)
            data_io.write_D(case_dir, world.rank, ts, cbs, q, prim, P.num_fluids, prim_wrt)
        io_time += time.perf_counter() - t
        n_saves += 1

    if cfl_dt:
        t_step, mytime, dt_init = 0, t_save_time * n_start, 0.0
        while mytime < t_stop_time:
            if (cfl_const and t_step == 0) or cfl_adap:
                dt = cfl_target * state.reduce(1)            # s_compute_dt: cfl_target * min over cells and ranks
            if t_step == 0:
                dt_init = dt
            if cfl_adap and dt < 1.0e-3 * dt_init:           # m_start_up.fpp:579-582
                raise FloatingPointError(f"Delta t = {dt}: Delta t has become too small")
            if (mytime + dt) >= t_stop_time:
                dt = t_stop_time - mytime
            if world.rank == 0 and t_step % t_print == 0:
                eta = wall_avg * (t_stop_time - mytime) / max(dt, sys.float_info.min)
                log(" [%3d%%] Time %16.6E dt = %16.6E @ Time Step = %8d Time Avg = %16.6E Time/step = %12.6E ETA (HH:MM:SS) = %d:%02d:%02d"
                    % (int(math.ceil(100.0 * (mytime / t_stop_time))), mytime, dt, t_step, wall_avg, wall, int(eta) // 3600,
                       (int(eta) % 3600) // 60, int(eta) % 60))
            if rti is not None or (run_time_info and world.size > 1):
                icfl = state.icfl_max(dt)
                if rti is not None:
                    rti.write(t_step, dt, mytime, icfl)
            ws = time.perf_counter()
            if hasattr(state, "set_time"):
                state.set_time(mytime)
            state.step(dt)
            wall = time.perf_counter() - ws
            if timed_api:
                if t_step < N_WARMUP:
                    state.stage_time_min = 0.0
                time_avg = state.stage_time_min
            wall_avg = (wall + (t_step - 2) * wall_avg) / (t_step - 1) if t_step >= 2 else 0.0
            mytime = mytime + dt
            t_step += 1
            if abs(math.fmod(mytime, t_save_time)) < dt or mytime >= t_stop_time:   # p_main.fpp:79-81
                save(int(mytime / t_save_time))                                      # save_count, m_start_up.fpp:758-759
        t1 = t_step

    while not cfl_dt and t_step != t1:
        # how many steps until the next event (print / save / end); whole segments go to the device at once
        # when nothing has to be observed in between
        if (mytime + dt) >= finaltime:              # m_start_up.fpp:589-594
            dt = finaltime - mytime
        if world.rank == 0 and (t_step - t0) % t_print == 0:
            log(_progress(t_step, t0, t1, wall_avg, wall))
        if rti is not None or (run_time_info and world.size > 1):
            icfl = state.icfl_max(dt)               # stage-1 primitives = the state at the start of the step
            if rti is not None:
                rti.write(t_step, dt, mytime, icfl)
        nseg = 1
        k0 = t_step - t0
        observed = run_time_info or k0 < N_WARMUP + N_TIMED or not hasattr(state, "steps")
        if not observed:
            nxt = min(t1 - 1, _next_multiple(t_step, t0, t_save), _next_multiple(t_step, t0, t_print))
            want, nseg, tt = max(1, nxt - t_step), 0, mytime
            while nseg < want and (tt + dt) < finaltime:   # a step whose dt gets clipped is always taken alone
                tt = tt + dt
                nseg += 1
            nseg = max(1, nseg)
        ws = time.perf_counter()
        if hasattr(state, "set_time"):
            state.set_time(mytime)   # GPU_UPDATE(device=[mytime]); steps advance it by dt themselves
        if nseg == 1:
            state.step(dt)
            if timed_api:
                if k0 < N_WARMUP:
                    state.stage_time_min = 0.0      # time_avg = 0 during the warm-up steps
                time_avg = state.stage_time_min
        else:
            state.steps(dt, nseg)
            state.sync()
        wall = (time.perf_counter() - ws) / nseg
        for _ in range(nseg):
            mytime = mytime + dt
            t_step += 1
            k = t_step - 1 - t0                     # m_time_steppers.fpp:604-608 (wall_time_avg over steps >= 2)
            wall_avg = (wall + (k - 2) * wall_avg) / (k - 1) if k >= 2 else 0.0
        if (t_step - t0) % t_save == 0 or t_step == t1:
            save(t_step)

    # ---- s_save_performance_metrics
    ns = P.time_stepper
    if not timed_api or time_avg <= 0.0:
        time_avg = wall_avg / ns if wall_avg > 0 else wall / ns
    time_final = world.max(time_avg)
    io_final = world.max(io_time / max(n_saves, 1))
    m_glb, n_glb, p_glb = blk.glb
    grind = time_final * 1.0e9 / (float(P.sys_size) * float(max(1, m_glb)) * float(max(1, n_glb)) * float(max(1, p_glb)))
    if world.rank == 0:
        log(" Performance: %24.16f ns/gp/eq/rhs" % grind)
        data_io.append_time_data(case_dir, world.size, time_final, grind, io_final)
    if rti is not None:
        rti.close(time.perf_counter() - run_start)
    out = {"grind_time": grind, "stage_time": time_final, "cons": state.get_cons(), "block": blk, "t_step": t_step,
           "kernel_launches": getattr(state, "kernel_launches", None)}
    if hasattr(state, "close"):
        state.close()
    world.barrier()
    return out


def _next_multiple(t_step, t0, every):
    """The first step index after ``t_step`` that is a multiple of ``every`` counted from ``t0``."""
    k = (t_step - t0) // every + 1
    return t0 + k * every


def _current_prim(state):
    """Primitive variables of the *current* conservative state (the writer converts before output, m_data_output.fpp:399-403)."""
    if hasattr(state, "reduce"):
        state.reduce(2)        # runs cons -> prim on the current state
    elif hasattr(state, "cons2prim"):
        state.cons2prim()
    return state.get_prim()


def _ghost_widths(blk, glb_cbs, b):
    """Neighbour cell widths from the global grid (what s_mpi_sendrecv_grid_variable_buffer delivers, m_mpi_common.fpp:1745-1796)."""
    gws = []
    for d in range(3):
        if d >= blk.nd:
            gws.append(None)
            continue
        gds = glb_cbs[d][1:] - glb_cbs[d][:-1]
        nglb, st, s = blk.glb[d] + 1, blk.start[d], blk.s[d]
        gb = np.array([gds[(st - b + i) % nglb] for i in range(b)]) if blk.bc[d][0] >= 0 else None
        ge = np.array([gds[(st + s + 1 + i) % nglb] for i in range(b)]) if blk.bc[d][1] >= 0 else None
        gws.append((gb, ge))
    return gws


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m mfc_b200", description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("run", help="run a reference case file (pre_process + simulation)")
    r.add_argument("input", help="case.py / case.json / simulation.inp")
    r.add_argument("-t", "--targets", nargs="+", default=["pre_process", "simulation"], choices=["pre_process", "simulation"])
    r.add_argument("--case-dir", default=None, help="where p_all/, D/, restart_data/ live (default: the case file's directory)")
    argv = list(sys.argv[1:] if argv is None else argv)
    extra = []
    if "--" in argv:   # arguments after -- are handed to the case file (the reference: ./mfc.sh run case.py -- --gbpp 16)
        i = argv.index("--")
        argv, extra = argv[:i], argv[i + 1:]
    a = ap.parse_args(argv)
    world = World()
    case = world.bcast(case_io.load_case(a.input, extra) if world.rank == 0 else None)
    case_dir = a.case_dir or os.path.dirname(os.path.abspath(a.input))
    if world.rank == 0:
        os.makedirs(case_dir, exist_ok=True)
        case_io.write_namelist(os.path.join(case_dir, "simulation.inp"), case)
    world.barrier()
    if "pre_process" in a.targets:
        run_pre_process(case, case_dir, world)
    if "simulation" in a.targets:
        run_simulation(case, case_dir, world)
    return 0


if __name__ == "__main__":
    sys.exit(main())
