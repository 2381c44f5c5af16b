#!/usr/bin/env python
"""bench.py -- grind time of the 5-equation RHS hot path on N B200s (driver contract).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libmfc5.so)
    python bench.py --impl reference --gpus N ...            # the CPU arm: the reference algorithm
                                                              # (oracle restatement, OpenMP, all host cores)

Workload (``config.workload``): ``examples/3D_performance_test`` -- the case the reference's published
figure of merit is quoted on (docs/documentation/expectedPerformance.md:92-118): 201^3 cells per GPU,
two fluids (8 equations), stretched grid, WENO5-M + HLLC + mpp_lim, TVD-RK3, BC -2/-6, FP64.  A "step" is
one s_tvd_rk call = 3 RK stages, each = cons->prim + ghost fill (+ halo exchange) + 3 fused sweeps + RK
update.  N > 1: weak scaling, 201^3 cells per GPU, block decomposition as the reference, halo faces over NCCL.

Metric: MFC's grind time, ns per grid point per equation per RHS evaluation,
``t_stage * 1e9 / (sys_size * m_glb * n_glb * p_glb)`` (m_start_up.fpp:674-677), with t_stage = the average
stage time over the K timed steps (the reference prints the minimum; that is reported as ``grind_min_stage``).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PUBLISHED_H100 = 0.38  # ns/gp/eq/rhs, H100 SXM5, BASELINE.md section 2 (other hardware; nothing published for B200)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--case", default="3D_performance_test", choices=["3D_performance_test", "5eq_rk3_weno3_hllc", "3D_shockdroplet", "2D_shockbubble", "1D_sodshocktube"])
    ap.add_argument("--m", type=int, default=200, help="cells-1 per direction per GPU (3D_performance_test)")
    ap.add_argument("--riemann", default="hllc", choices=["hllc", "hll"])
    ap.add_argument("--gbpp", type=int, default=16, help="5eq_rk3_weno3_hllc: the reference's --gbpp")
    ap.add_argument("--weno5z", action="store_true", help="3D_shockdroplet: the north star's WENO5-Z override")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE.json's other configurations (C1 .. C4, HLL variant, strong scaling)")
    ap.add_argument("--other-budget", type=float, default=150.0, help="seconds the other configurations may take in total")
    ap.add_argument("--cpu-steps", type=int, default=2)
    return ap.parse_args()


def most_cubic(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n) or (n, 1, 1)


def build_case(args, nranks):
    """Global case dictionary + processor grid."""
    from mfc_b200 import cases, decomp
    if args.case == "3D_performance_test":
        case = cases.perf_test_3d(m=args.m, riemann=args.riemann)
        name = f"examples/3D_performance_test {args.m + 1}^3 cells/GPU, WENO5-M, {args.riemann.upper()}, mpp_lim, RK3"
    elif args.case == "5eq_rk3_weno3_hllc":
        case = cases.bench_5eq_weno3_hllc(gbpp=args.gbpp, riemann=args.riemann)
        name = f"benchmarks/5eq_rk3_weno3_{args.riemann} --gbpp {args.gbpp}: {case['m'] + 1}x{case['n'] + 1}x{case['p'] + 1} cells/GPU, WENO3-M"
    elif args.case == "2D_shockbubble":
        case = cases.shockbubble_2d()
        name = f"examples/2D_shockbubble {case['m'] + 1}x{case['n'] + 1} cells, air/helium, WENO5-M, HLLC"
    elif args.case == "1D_sodshocktube":
        case = cases.sod_1d()
        name = f"examples/1D_sodshocktube {case['m'] + 1} cells, WENO5-M, HLLC"
    else:
        case = cases.shockdroplet_3d(**(dict(weno_order=5, mapped_weno="F", wenoz="T") if args.weno5z else {}))
        name = "examples/3D_shockdroplet 201x101x101 " + ("WENO5-Z override" if args.weno5z else "WENO3-M as shipped") + ", HLLC"
    dims = (1, 1, 1)
    if nranks > 1:
        if args.scaling == "weak":
            dims = most_cubic(nranks)
            for d, x in enumerate(("m", "n", "p")):
                case[x] = (case[x] + 1) * dims[d] - 1
            assert decomp.factorize(nranks, case["m"], case["n"], case["p"], case["weno_order"]) == dims, "processor grid differs from the reference's rule"
        else:
            dims = decomp.factorize(nranks, case["m"], case["n"], case["p"], case["weno_order"])
    return case, dims, name


def setup_rank(case, rank, nranks, dims):
    """Local parameters, grid and initial condition of one rank."""
    from mfc_b200 import decomp, pre_process
    from mfc_b200.params import params_from_case
    blk = decomp.Block(case, rank, nranks, dims)
    lcase = blk.local_case(case)
    lcase["parallel_io"] = "T"  # local grid = slice of the global grid
    P = params_from_case(lcase)
    cbs, ccs, gws = blk.grid(dict(case, parallel_io="T"), P.buff_size)
    _, cons = pre_process.initial_condition(lcase, ccs, ds_min=pre_process.grid_minima(case))
    return blk, lcase, P, cbs, gws, cons


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1, busy0=None):
        """Clocks over the timed region [t0, t1]; if nvidia-smi delivered fewer than 3 samples there, over
        the whole GPU-busy window since `busy0` (warm-up + timed steps), and the window is named."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out = self._window(t0, t1)
        out["window"] = "timed region"
        if out["samples"] < 3 and busy0 is not None:
            out = self._window(busy0, t1)
            out["window"] = "warm-up + timed region (timed region shorter than 3 nvidia-smi samples)"
        return out

    def _window(self, t0, t1):
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[0]))
                    pw.append(float(f[2]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def config_dict(name, ncell_loc, ncell_glb, N, dims, world):
    """The `config` object, identical for both arms (the driver compares them key by key)."""
    return {"workload": name, "cells_per_gpu": ncell_loc, "cells_global": ncell_glb, "sys_size": N, "proc_grid": list(dims),
            "l2": "inputs larger than L2 (state 4 x %.0f MB per GPU)" % (N * ncell_loc * 8 / 1e6),
            "vs_baseline_ref": "0.38 ns/gp/eq/rhs published for H100 SXM5 (BASELINE.md); nothing published for B200",
            "parallelism": "block domain decomposition %dx%dx%d, NCCL halo" % tuple(dims) if world > 1 else "single GPU"}


def reference_toolchain():
    """the reference's own CPU build needs gfortran (or nvfortran), an MPI, fypp and cmake.  Probe the box
    and say what is there; without them the CPU arm is the C restatement (`kind: "port"`)."""
    import shutil
    tools = {t: shutil.which(t) for t in ("gfortran", "nvfortran", "mpirun", "mpif90", "fypp", "cmake")}
    have = bool((tools["gfortran"] or tools["nvfortran"]) and tools["mpirun"] and tools["fypp"] and tools["cmake"])
    return {"found": {k: v for k, v in tools.items() if v}, "missing": sorted(k for k, v in tools.items() if not v),
            "mfc_cpu_build_possible": have}


def multi_gpu_parity(args, rank, world, local_rank):
    """Driver-visible multi-GPU result check (run before the timed region): 5 RK3 steps of the bench case at 60^3 cells per GPU,
    decomposed over `world` GPUs with NCCL halo exchange, gathered to rank 0 and compared with rank 0's own single-GPU run of
    the same global case; bar 1e-12 of max|q| per variable (the halo values are copies; only the ghost cell-boundary coordinates
    are rebuilt per rank, an ulp-level change of the WENO tables on the stretched grid).  Reference: s_mpi_sendrecv_variables_buffers
    m_mpi_common.fpp:620-1065."""
    import torch
    import torch.distributed as dist
    from mfc_b200 import decomp, lib
    from mfc_b200.lib import Mfc5
    sargs = argparse.Namespace(**vars(args))
    sargs.case, sargs.m, sargs.scaling = "3D_performance_test", 59, "weak"
    case, dims, _ = build_case(sargs, world)
    nsteps, dt = 5, None
    blk, lcase, P, cbs, gws, cons = setup_rank(case, rank, world, dims)
    dt = float(case["dt"])
    P.device = local_rank
    g = Mfc5(P, cbs, ghost_widths=gws)
    uid = [lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(uid[0], rank, world)
    g.set_cons(cons)
    g.steps(dt, nsteps)
    g.sync()
    loc = torch.from_numpy(g.get_cons()).cuda()
    parts = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc)
    launches = g.kernel_launches
    g.close()
    if rank != 0:
        return None
    _, _, Pg, cbg, _, cong = setup_rank(case, 0, 1, (1, 1, 1))
    Pg.device = local_rank
    s = Mfc5(Pg, cbg)
    s.set_cons(cong)
    s.steps(dt, nsteps)
    s.sync()
    full = s.get_cons()
    s.close()
    worst = 0.0
    for r in range(world):
        ref = decomp.Block(case, r, world, dims).slab(full)
        a = parts[r].cpu().numpy()
        for v in range(a.shape[0]):
            worst = max(worst, float(np.abs(a[v] - ref[v]).max() / (np.abs(full[v]).max() + 1e-300)))
    moved = float(np.abs(full - cong).max() / np.abs(cong).max())
    return {"what": "decomposed run == single-GPU run of the same global case, 3D_performance_test at 60^3 cells per GPU, %d RK3 steps" % nsteps,
            "cells_global": int((case["m"] + 1) * (case["n"] + 1) * (case["p"] + 1)), "proc_grid": list(dims), "max_rel": worst, "tol": 1e-12,
            "ok": bool(worst <= 1e-12 and moved > 0.0), "state_change": moved, "kernel_launches_per_rank": int(launches)}


def time_config(sargs, rank, world, local_rank, steps, warmup, peak):
    """Device-resident time per RK3 step of one more configuration (same timing rules as the headline: warm-up, CUDA events on
    the library's compute stream, barrier + synchronise on both sides, max over ranks).  Collective: every rank calls it."""
    import torch
    import torch.distributed as dist
    from mfc_b200 import lib
    from mfc_b200.lib import Mfc5
    t_setup = time.time()
    case, dims, name = build_case(sargs, world)
    blk, lcase, P, cbs, gws, cons = setup_rank(case, rank, world, dims)
    P.device = local_rank
    g = Mfc5(P, cbs, ghost_widths=gws)
    if world > 1:
        uid = [lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
    g.set_cons(cons)
    del cons
    stream = torch.cuda.ExternalStream(g.compute_stream, device=torch.device("cuda", local_rank))
    dt = float(case["dt"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.sync()
    barrier()
    g.steps(dt, warmup)
    barrier()
    l0 = g.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.steps(dt, steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    launches = g.kernel_launches - l0
    nan_count = g.reduce(2)
    graph_state = g.graph_state
    N, ns, nd = g.N, P.time_stepper, P.num_dims
    g.close()
    ms_step = ms / steps
    ncell_loc = (P.m + 1) * (P.n + 1) * (P.p + 1)
    ncell_glb = (case["m"] + 1) * (case["n"] + 1) * (case["p"] + 1)
    stage_bytes = (12 if nd == 3 else 9 if nd == 2 else 6) * N * 8
    ach = stage_bytes * ncell_loc / (ms_step * 1e-3 / ns) / 1e9
    return {"workload": name, "cells_per_gpu": ncell_loc, "cells_global": ncell_glb, "sys_size": N, "proc_grid": list(dims), "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "value": grind(ms_step * 1e-3 / ns, (case["m"], case["n"], case["p"]), N),
            "unit": "ns/gp/eq/rhs", "cells_steps_per_s": ncell_glb / (ms_step * 1e-3),
            "roofline_stage": {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_cell_per_stage": stage_bytes},
            "gpu_launches": int(launches), "graph_replay": graph_state, "nan_count": nan_count, "setup_s": round(time.time() - t_setup - ms * 1e-3, 1)}


def other_configs(args, rank, world, local_rank, peak, budget_s):
    """BASELINE.json's other configurations, timed device-resident after the headline measurement (parity for each of them is the
    test suite's job: tests/test_gpu_parity.py).  N = 1: C1 .. C4 and the HLL variant of C5; N > 1: the strong-scaling case
    (C4, fixed global size over N GPUs) and the HLL variant of the weak-scaling case."""
    mk = lambda **kw: argparse.Namespace(**dict(vars(args), **kw))
    if world == 1:
        plan = [("C5_hll", mk(case="3D_performance_test", riemann="hll", scaling="weak"), 10, 3),
                ("C4_shockdroplet_weno5z", mk(case="3D_shockdroplet", weno5z=True, riemann="hllc"), 50, 5),
                ("C4_shockdroplet_as_shipped", mk(case="3D_shockdroplet", weno5z=False, riemann="hllc"), 50, 5),
                ("C2_shockbubble_2d", mk(case="2D_shockbubble", riemann="hllc"), 100, 10),
                ("C1_sodshocktube_1d", mk(case="1D_sodshocktube", riemann="hllc"), 200, 20),
                ("C3_5eq_rk3_weno3_hllc_gbpp180", mk(case="5eq_rk3_weno3_hllc", gbpp=180, riemann="hllc"), 5, 3)]
    else:
        plan = [("C5_hll_weak", mk(case="3D_performance_test", riemann="hll", scaling="weak"), 10, 3),
                ("C4_strong_weno5z", mk(case="3D_shockdroplet", weno5z=True, riemann="hllc", scaling="strong"), 100, 10),
                ("C4_strong_as_shipped", mk(case="3D_shockdroplet", weno5z=False, riemann="hllc", scaling="strong"), 100, 10)]
    out, t0 = {}, time.time()
    for key, sargs, steps, warm in plan:
        import torch
        flag = torch.tensor([1.0 if time.time() - t0 < budget_s else 0.0], device="cuda")
        if world > 1:
            import torch.distributed as dist
            dist.broadcast(flag, src=0)   # one decision for all ranks
        if flag.item() == 0.0:
            out[key] = {"skipped": "time budget of the default run (%d s) used up" % budget_s}
            continue
        out[key] = time_config(sargs, rank, world, local_rank, steps, warm, peak)
    return out


def cpu_arm(case, P, cbs, cons, nsteps, warm):
    """Time the oracle (reference algorithm restated in C, OpenMP) on this box's host cores."""
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1); must be set before libgomp starts
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle(P, cbs, omp=True)
    o.set_cons(cons)
    dt = float(case["dt"])
    for _ in range(warm):
        o.step(dt)
    ts = []
    for _ in range(nsteps):
        t = time.perf_counter()
        o.step(dt)
        ts.append(time.perf_counter() - t)
    return ts, o.L.orc_num_threads(), o


def grind(t_stage_s, P_glb_mnp, sys_size):
    m, n, p = P_glb_mnp
    return t_stage_s * 1e9 / (sys_size * float(m) * float(max(n, 1)) * float(max(p, 1)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    # the workload named in `config` is the native arm's (N GPUs: N x the per-GPU block); the CPU arm times a bounded sample of
    # it -- one per-GPU block, or a 100^3 cube of it -- and reports the intensive metric (ns per cell per equation per RHS)
    gcase, gdims, gname = build_case(args, max(1, args.gpus))
    case, dims, name = build_case(args, 1)
    full_cells = (case["m"] + 1) * (case["n"] + 1) * (case["p"] + 1)
    sample = "one RK3 step of one GPU's block of the workload (%d cells) per timed step" % full_cells
    blk, lcase, P, cbs, gws, cons = setup_rank(case, 0, 1, (1, 1, 1))
    t = time.perf_counter()
    ts, threads, o = cpu_arm(lcase, P, cbs, cons, 1, 0)
    est = ts[0] * (args.steps + args.warmup)
    if est > 240.0 and args.case == "3D_performance_test":
        # bounded sample: a smaller cube of the same case (same scheme, same per-cell arithmetic)
        del o
        msmall = 99
        sargs = argparse.Namespace(**vars(args))
        sargs.m = msmall
        case, dims, _ = build_case(sargs, 1)
        blk, lcase, P, cbs, gws, cons = setup_rank(case, 0, 1, (1, 1, 1))
        sample = f"one RK3 step of a {msmall + 1}^3 cube of the same case per timed step (one GPU's block projected at {est:.0f} s for the run)"
    else:
        del o
    ts, threads, _ = cpu_arm(lcase, P, cbs, cons, args.steps, args.warmup)
    t_step = float(np.mean(ts))
    ns = P.time_stepper
    g = grind(t_step / ns, (case["m"], case["n"], case["p"]), P.sys_size)
    ncell = (case["m"] + 1) * (case["n"] + 1) * (case["p"] + 1)
    out = {
        "impl": "reference", "metric": "grind_time", "value": g, "unit": "ns/gp/eq/rhs", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": False, "scaling": args.scaling, "vs_baseline": g / PUBLISHED_H100,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(gname, full_cells, (gcase["m"] + 1) * (gcase["n"] + 1) * (gcase["p"] + 1), P.sys_size, gdims, max(1, args.gpus)),
        "cells_steps_per_s": ncell / t_step, "cells_timed": ncell,
        "note": "reference algorithm as the CPU oracle restatement (kind: port): MFC's Fortran cannot be built in this image",
        "cpu_baseline": {"value": g, "unit": "ns/gp/eq/rhs", "cores": threads, "kind": "port", "sample": sample,
                         "reference_toolchain": reference_toolchain()},
        "e2e": {"value": g, "unit": "ns/gp/eq/rhs", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the native arm)")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mfc_b200 import lib
    from mfc_b200.lib import Mfc5

    parity = None
    if world > 1:
        parity = multi_gpu_parity(args, rank, world, local_rank)
        ok = [parity["ok"] if rank == 0 else None]
        dist.broadcast_object_list(ok, src=0)
        if not ok[0]:
            if rank == 0:
                print(json.dumps({"metric": "grind_time", "n_gpus": world, "error": "multi-GPU parity check failed", "parity_check": parity}), flush=True)
            dist.destroy_process_group()
            raise SystemExit(3)
    case, dims, name = build_case(args, world)
    blk, lcase, P, cbs, gws, cons = setup_rank(case, rank, world, dims)
    P.device = local_rank
    g = Mfc5(P, cbs, ghost_widths=gws)
    if world > 1:
        uid = [lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
    N, ns = g.N, P.time_stepper
    dt = float(case["dt"])
    ncell_loc = (P.m + 1) * (P.n + 1) * (P.p + 1)
    ncell_glb = (case["m"] + 1) * (case["n"] + 1) * (case["p"] + 1)

    # pinned host state (the caller's q_cons_vf arrays); with several ranks on the node each one first moves next to its GPU,
    # as the reference's MPI launcher binds its ranks (mfc_b200/affinity.py)
    affinity = None
    if world > 1 and not os.environ.get("MFC5_NO_BIND"):
        from mfc_b200.affinity import bind_to_gpu
        affinity = bind_to_gpu(local_rank, world)
    host = torch.empty(g.shape, dtype=torch.float64, pin_memory=True)
    hq = host.numpy()
    hq[...] = cons
    g.set_cons(hq)
    stream = torch.cuda.ExternalStream(g.compute_stream, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.sync()

    # ---- device-resident throughput: W warm-up steps, K timed steps, CUDA events on the compute stream
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    barrier()
    busy0 = time.time()
    g.steps(dt, args.warmup)
    barrier()
    l0 = g.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    g.steps(dt, args.steps)
    e1.record(stream)
    barrier()
    t1 = time.time()
    launches = g.kernel_launches - l0
    graph_state = g.graph_state   # 1: the timed steps were CUDA-graph replays; -1: capture failed, eager fallback
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1, busy0) if sampler else None
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_step = ms / args.steps
    nan_count = g.reduce(2)

    # ---- per-kernel-class device times (CUDA events inside the library), 3 profiled steps
    # (median over the profiled steps per class, scaled back to a sum: one disturbed launch -- seen once as a 1.09 ms
    # y sweep average against 0.79 ms in every other run and under ncu -- must not pick the "dominant" kernel)
    prof = {}
    samples = {}
    nprof = 5
    g.stage_time_min = 0.0
    for _ in range(nprof):
        g.step(dt)
        for k, (kms, nl) in g.last_kernel_ms().items():
            samples.setdefault(k, []).append(kms)
            a = prof.setdefault(k, [0.0, 0])
            a[1] += nl
    for k, v in samples.items():
        prof[k][0] = sorted(v)[len(v) // 2] * len(v)
    stage_min_s = g.stage_time_min
    if world > 1:
        tt = torch.tensor([stage_min_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stage_min_s = float(tt.item())

    # ---- end to end through the C ABI with host buffers: upload q_cons, one s_tvd_rk, download q_cons
    e2e = None
    if not args.no_e2e:
        ne = max(3, min(args.steps, 5))
        for _ in range(2):
            g.set_cons(hq); g.steps(dt, 1); g.L.mfc5_download_cons(g.h, g._field_ptrs(hq), 0)
        barrier()
        t = time.perf_counter()
        for _ in range(ne):
            g.set_cons(hq)
            g.steps(dt, 1)
            rc = g.L.mfc5_download_cons(g.h, g._field_ptrs(hq), 0)
            assert rc == 0
        barrier()
        te = (time.perf_counter() - t) / ne
        if world > 1:
            tt = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        # the same calls with the cadence the case file itself asks for: the state crosses PCIe once per t_step_save steps
        # (examples/3D_performance_test/case.py: t_step_save = 30), i.e. how the reference's own time loop uses s_tvd_rk
        nsave = int(case.get("t_step_save", 30) or 30)
        nsave = max(1, min(nsave, 30))
        barrier()
        t = time.perf_counter()
        g.set_cons(hq)
        g.steps(dt, nsave)
        rc = g.L.mfc5_download_cons(g.h, g._field_ptrs(hq), 0)
        assert rc == 0
        barrier()
        ta = (time.perf_counter() - t) / nsave
        if world > 1:
            tt = torch.tensor([ta], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ta = float(tt.item())
        nbytes = int(hq.nbytes)
        e2e = {"value": grind(te / ns, (case["m"], case["n"], case["p"]), N), "unit": "ns/gp/eq/rhs", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "ms_per_step": te * 1e3, "timer": "host wall clock around upload + s_tvd_rk + download, synchronised both sides",
               "note": "worst case: the whole conservative state crosses PCIe both ways every step; the reference (and this library) keep it resident between saves",
               "pcie_gbs_per_direction": 2 * nbytes / max(te - ms_step * 1e-3, 1e-9) / 1e9 / 2, "rank_affinity": affinity,
               "at_case_save_cadence": {"steps_per_save": nsave, "ms_per_step": ta * 1e3,
                                        "value": grind(ta / ns, (case["m"], case["n"], case["p"]), N), "unit": "ns/gp/eq/rhs"}}

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s"

    # ---- BASELINE.json's other configurations (collective: every rank takes part)
    others = None
    if not args.no_other_configs and args.case == "3D_performance_test" and args.riemann == "hllc" and args.scaling == "weak":
        others = other_configs(args, rank, world, local_rank, peak, args.other_budget)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md 8d / DESIGN.md)
    nd = P.num_dims
    last = ["sweep_x", "sweep_y", "sweep_z"][nd - 1]
    alg = {"cons2prim": 2 * N * 8, "sweep_x": 2 * N * 8, "sweep_y": 3 * N * 8, "sweep_z": 3 * N * 8}
    alg[last] = (5 if nd > 1 else 4) * N * 8  # prim + rhs + q1 + q2 read, q written (1-D: no rhs read)
    fused_c2p = "cons2prim" not in prof or prof["cons2prim"][1] == 0
    if fused_c2p:
        # case-optimised path: the RK-fused last sweep also writes the next stage's primitive variables (over the rhs entries it
        # has consumed), which replaces the separate cons -> prim pass (N read + N written): one more N written here
        alg[last] += N * 8
    kern = {}
    for k, (kms, nl) in prof.items():
        if nl == 0:
            continue
        per = kms / nl
        kern[k] = {"ms_per_launch": per, "launches_per_step": nl / nprof, "ms_per_step": kms / nprof}
        if k in alg:
            # one pass over all cells of the rank = the launches of this class in one stage (the x sweep is split into
            # interior + edge launches when a halo exchange is in flight)
            kern[k]["ms_per_pass"] = kms / nprof / ns
            kern[k]["achieved_gbs"] = alg[k] * ncell_loc / (kern[k]["ms_per_pass"] * 1e-3) / 1e9
    dom = max((k for k in kern if k in alg), key=lambda k: kern[k]["ms_per_step"])
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_cell": alg[dom], "share_of_step": kern[dom]["ms_per_step"] / sum(v["ms_per_step"] for v in kern.values())}
    if fused_c2p and dom == last:
        s5 = (5 if nd > 1 else 4) * N * 8
        roof["algorithmic_bytes_note"] = ("%d B/cell = SURVEY 8(d)'s last sweep with RK (%d B) + the N doubles of the cons->prim pass's output, which this "
                                          "kernel writes instead of a separate pass (whose N doubles of input are no longer read at all)" % (alg[dom], s5))
        roof["frac_at_survey_bytes"] = s5 * ncell_loc / (kern[dom]["ms_per_pass"] * 1e-3) / 1e9 / peak
    stage_bytes = (12 if nd == 3 else 9 if nd == 2 else 6) * N * 8
    t_stage = ms_step * 1e-3 / ns
    roof_stage = {"achieved": stage_bytes * ncell_loc / t_stage / 1e9, "peak": peak, "unit": "GB/s",
                  "frac": stage_bytes * ncell_loc / t_stage / 1e9 / peak, "algorithmic_bytes_per_cell_per_stage": stage_bytes}
    prof_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof_file):
        try:
            with open(prof_file) as f:
                tj = json.load(f)
            roof["traffic"] = tj.get(dom)
            roof["traffic_source"] = tj.get("_source", "committed ncu capture (profiles/traffic.json), not measured in this run")
        except (OSError, ValueError):
            pass

    # the sweeps are FP64-pipe bound: the same launch times against the measured FP64 issue peak
    fp64 = None
    cnt_file = os.path.join(ROOT, "profiles", "fp64_counts.json")
    if args.case == "3D_performance_test" and args.riemann == "hllc" and os.path.exists(cnt_file):
        try:
            with open(cnt_file) as f:
                cnt = json.load(f)
            sm_hz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
            pk = cnt["peak_thread_inst_per_clk_per_sm"] * 148 * sm_hz
            per = {k: cnt["thread_inst_per_cell"][k] * ncell_loc / (kern[k]["ms_per_pass"] * 1e-3) for k in cnt["thread_inst_per_cell"] if k in kern}
            tot_inst = sum(cnt["thread_inst_per_cell"][k] for k in per) * ncell_loc
            tot_t = sum(kern[k]["ms_per_pass"] for k in per) * 1e-3
            fp64 = {"unit": "FP64 thread instructions/s (DFMA+DMUL+DADD)", "peak": pk, "achieved": tot_inst / tot_t, "frac": tot_inst / tot_t / pk,
                    "per_sweep_frac": {k: v / pk for k, v in per.items()}, "thread_inst_per_cell": cnt["thread_inst_per_cell"],
                    "stage_floor_ms": tot_inst / pk * 1e3, "source": "profiles/fp64_counts.json (ncu) + profiles/fp64_peak_r01.txt"}
        except (OSError, KeyError, ValueError, ZeroDivisionError):
            fp64 = None

    value = grind(t_stage, (case["m"], case["n"], case["p"]), N)
    out = {
        "metric": "grind_time", "value": value, "unit": "ns/gp/eq/rhs", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": False, "scaling": args.scaling, "vs_baseline": value / PUBLISHED_H100,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(name, ncell_loc, ncell_glb, N, dims, world),
        "cells_steps_per_s": ncell_glb / (ms_step * 1e-3),
        "grind_min_stage": grind(stage_min_s, (case["m"], case["n"], case["p"]), N),
        "clocks": clocks, "gpu_launches": int(launches), "nan_count": nan_count,
        "roofline": roof, "roofline_stage": roof_stage, "kernels": kern,
    }
    if parity is not None:
        out["parity_check"] = parity
    if others is not None:
        out["other_configs"] = others
    out["graph_replay"] = graph_state
    if fp64 is not None:
        out["roofline_fp64"] = fp64
    if e2e is not None:
        out["e2e"] = e2e
    # ---- CPU baseline on this box's host cores (bounded sample), N = 1 only
    if world == 1 and not args.no_cpu_baseline:
        g.close()
        del g
        bm = 200 if (args.case == "3D_performance_test" and ncell_loc > 12_000_000) else None
        if bm is not None:
            sargs = argparse.Namespace(**vars(args))
            sargs.m = bm
            bcase, _, _ = build_case(sargs, 1)
            _, blcase, bP, bcbs, _, bcons = setup_rank(bcase, 0, 1, (1, 1, 1))
            sample = f"{args.cpu_steps} RK3 steps (after 1 warm-up) of a {bm + 1}^3-cell cube of the same case"
        else:
            bcase, blcase, bP, bcbs, bcons = case, lcase, P, cbs, cons
            sample = f"{args.cpu_steps} RK3 steps (after 1 warm-up) of the full workload"
        ts, threads, _ = cpu_arm(blcase, bP, bcbs, bcons, args.cpu_steps, 1)
        gcpu = grind(float(np.mean(ts)) / ns, (bcase["m"], bcase["n"], bcase["p"]), N)
        out["cpu_baseline"] = {"value": gcpu, "unit": "ns/gp/eq/rhs", "cores": threads, "kind": "port", "sample": sample,
                               "reference_toolchain": reference_toolchain()}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
