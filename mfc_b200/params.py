"""ctypes mirror of ``include/mfc5_params.h`` and the case.py -> ``mfc5_params`` translation.

The keys accepted are the reference's own namelist names (``m``, ``weno_order``, ``bc_x%beg``,
``fluid_pp(1)%gamma`` ...; reference ``src/simulation/m_start_up.fpp:81-130``), with defaults as in
``src/simulation/m_global_parameters.fpp:330-420`` and named enums as in
``toolchain/mfc/params/definitions.py:299-359``.  Parameters that would leave the 5-equation
inviscid envelope are rejected.
"""
import ctypes as C

MAX_FLUIDS = 4

_NAMED = {
    "model_eqns": {"gamma_law": 1, "5eq": 2, "6eq": 3},
    "riemann_solver": {"hll": 1, "hllc": 2, "hlld": 4, "lax_friedrichs": 5},
    "time_stepper": {"rk1": 1, "rk2": 2, "rk3": 3},
    "wave_speeds": {"direct": 1, "pressure": 2},
    "avg_state": {"roe": 1, "arithmetic": 2},
    "recon_type": {"weno": 1, "muscl": 2},
    "muscl_order": {"first_order": 1, "second_order": 2},
    "muscl_lim": {"unlimited": 0, "minmod": 1, "mc": 2, "van_albada": 3, "van_leer": 4, "superbee": 5},
    "int_comp": {"off": 0, "thinc": 1, "mthinc": 2},
}


class Mfc5Params(C.Structure):
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("p", C.c_int),
        ("num_fluids", C.c_int), ("model_eqns", C.c_int),
        ("weno_order", C.c_int), ("mapped_weno", C.c_int), ("wenoz", C.c_int), ("null_weights", C.c_int),
        ("weno_eps", C.c_double),
        ("riemann_solver", C.c_int), ("hll_u_interface", C.c_int), ("avg_state", C.c_int),
        ("wave_speeds", C.c_int), ("mpp_lim", C.c_int), ("mixture_err", C.c_int),
        ("time_stepper", C.c_int),
        ("dt", C.c_double),
        ("bc", (C.c_int * 2) * 3),
        ("bc_vb", (C.c_double * 3) * 3), ("bc_ve", (C.c_double * 3) * 3),
        ("grcbc_in", C.c_int * 3), ("grcbc_out", C.c_int * 3), ("grcbc_vel_out", C.c_int * 3),
        ("vel_in", (C.c_double * 3) * 3), ("vel_out", (C.c_double * 3) * 3),
        ("pres_in", C.c_double * 3), ("pres_out", C.c_double * 3),
        ("alpha_rho_in", (C.c_double * MAX_FLUIDS) * 3), ("alpha_in", (C.c_double * MAX_FLUIDS) * 3),
        ("fluid_gamma", C.c_double * MAX_FLUIDS), ("fluid_pi_inf", C.c_double * MAX_FLUIDS),
        ("fluid_qv", C.c_double * MAX_FLUIDS),
        ("device", C.c_int), ("fast_math", C.c_int), ("low_Mach", C.c_int), ("alt_soundspeed", C.c_int), ("muscl_order", C.c_int), ("muscl_lim", C.c_int), ("muscl_eps", C.c_double),
        ("teno", C.c_int), ("mp_weno", C.c_int), ("teno_CT", C.c_double),
        ("bf", C.c_int * 3), ("reserved3", C.c_int),
        ("bf_k", C.c_double * 3), ("bf_w", C.c_double * 3), ("bf_p", C.c_double * 3), ("bf_g", C.c_double * 3),
        ("int_comp", C.c_int), ("reserved4", C.c_int), ("ic_eps", C.c_double), ("ic_beta", C.c_double),
        ("wenoz_q", C.c_double),
        ("viscous", C.c_int), ("weno_Re_flux", C.c_int), ("weno_avg", C.c_int), ("reserved5", C.c_int),
        ("fluid_Re", (C.c_double * 2) * MAX_FLUIDS),
    ]

    @property
    def num_dims(self):
        return 1 + (self.n > 0) + (self.p > 0)

    @property
    def sys_size(self):
        return 2 * self.num_fluids + self.num_dims + 1

    @property
    def buff_size(self):
        # src/common/m_helper_basic.fpp:127-135 (inviscid WENO: weno_polyn + 2; MUSCL: muscl_polyn + 2 = muscl_order + 2)
        if self.muscl_order > 0:
            return self.muscl_order + 2
        if self.viscous:   # 2 * weno_polyn + 2
            return (self.weno_order - 1) + 2
        return (self.weno_order - 1) // 2 + 2

    @property
    def recon_order(self):
        return self.muscl_order if self.muscl_order > 0 else self.weno_order

    def copy(self):
        q = Mfc5Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Mfc5Params))
        return q


def _logical(v):
    if isinstance(v, str):
        return 1 if v.strip().upper() in ("T", ".TRUE.", "TRUE") else 0
    return 1 if v else 0


def _val(case, key, default=None):
    v = case.get(key, default)
    if isinstance(v, str) and key in _NAMED:
        v = _NAMED[key].get(v, v)
    return v


_UNSUPPORTED_TRUE = [
    "surface_tension", "bubbles_euler", "bubbles_lagrange", "qbmm", "hypoelasticity",
    "hyperelasticity", "mhd", "relativity", "chemistry", "igr", "ib", "cyl_coord", "acoustic_source", "relax",
    "cont_damage", "hyper_cleaning",
]


class UnsupportedCase(ValueError):
    """Raised (like the reference's ``@:PROHIBIT`` / ``s_mpi_abort``) for cases outside the envelope."""


def params_from_case(case, device=0, fast_math=1, oracle_only=False):
    """Translate a case dictionary (what a reference ``case.py`` prints) into ``Mfc5Params``.

    ``fast_math`` (product-side knob, ignored by the oracle): 1 = division-light arithmetic (default),
    0 = the reference's operation order."""
    for k in _UNSUPPORTED_TRUE:
        if _logical(case.get(k, "F")):
            raise UnsupportedCase(f"{k}=T is outside the 5-equation inviscid hot path")
    P = Mfc5Params()
    P.m, P.n, P.p = int(case["m"]), int(case.get("n", 0)), int(case.get("p", 0))
    P.num_fluids = int(case["num_fluids"])
    P.model_eqns = int(_val(case, "model_eqns"))
    if P.model_eqns != 2:
        raise UnsupportedCase("only model_eqns = 2 (5-equation) is on the hot path")
    if P.num_fluids > MAX_FLUIDS:
        raise UnsupportedCase(f"num_fluids > {MAX_FLUIDS}")
    if int(_val(case, "recon_type", 1)) == 2:  # MUSCL: m_muscl.fpp; defaults m_global_parameters.fpp:718-724
        P.muscl_order = int(_val(case, "muscl_order", 1))
        if P.muscl_order not in (1, 2):
            raise UnsupportedCase("muscl_order must be 1 or 2")
        P.muscl_lim = int(_val(case, "muscl_lim", 0) or 0)
        if P.muscl_lim not in (0, 1, 2, 3, 4, 5):
            raise UnsupportedCase("muscl_lim must be 0..5")
        P.muscl_eps = float(case["muscl_eps"]) if "muscl_eps" in case else (1e-9 if P.muscl_lim <= 2 else 1e-6)
        P.weno_order = 2 * P.muscl_order + 1   # same stencil window / buffer geometry as the WENO of that width
    else:
        P.weno_order = int(case["weno_order"])
        if P.weno_order not in (1, 3, 5, 7):   # toolchain/mfc/case_validator.py:401
            raise UnsupportedCase("weno_order must be 1, 3, 5, or 7")
    P.mapped_weno = _logical(case.get("mapped_weno", "F"))
    P.wenoz = _logical(case.get("wenoz", "F"))
    P.null_weights = _logical(case.get("null_weights", "F"))
    P.teno = _logical(case.get("teno", "F"))
    if P.teno:  # toolchain/mfc/case_validator.py:942-966
        if P.muscl_order > 0 or P.weno_order not in (5, 7):
            raise UnsupportedCase("teno requires weno_order = 5 or 7")
        if P.mapped_weno or P.wenoz:
            raise UnsupportedCase("Only one of mapped_weno, wenoz, or teno can be set to true")
        if "teno_CT" not in case or float(case["teno_CT"]) <= 0:
            raise UnsupportedCase("teno requires teno_CT to be set. A typical value is 1e-6")
        P.teno_CT = float(case["teno_CT"])
    P.wenoz_q = float(case.get("wenoz_q", 1.0) or 1.0)
    if P.wenoz and P.weno_order == 7:  # toolchain/mfc/case_validator.py:959-960
        if "wenoz_q" not in case:
            raise UnsupportedCase("wenoz at 7th order requires wenoz_q to be set (should be 2, 3, or 4)")
        if P.wenoz_q not in (2.0, 3.0, 4.0):
            raise UnsupportedCase("wenoz_q must be either 2, 3, or 4")
    P.mp_weno = _logical(case.get("mp_weno", "F"))
    if P.mp_weno and (P.muscl_order > 0 or P.weno_order != 5):  # toolchain/mfc/case_validator.py:967
        raise UnsupportedCase("mp_weno requires weno_order = 5")
    P.weno_eps = float(case.get("weno_eps", 1e-16))
    P.riemann_solver = int(_val(case, "riemann_solver"))
    if P.riemann_solver not in (1, 2, 5):
        raise UnsupportedCase("riemann_solver must be HLL (1), HLLC (2) or Lax-Friedrichs (5)")
    P.hll_u_interface = _logical(case.get("hll_u_interface", "F"))
    P.avg_state = int(_val(case, "avg_state", 2))
    P.wave_speeds = int(_val(case, "wave_speeds", 1))
    P.mpp_lim = _logical(case.get("mpp_lim", "F"))
    P.mixture_err = _logical(case.get("mixture_err", "F"))
    P.alt_soundspeed = _logical(case.get("alt_soundspeed", "F"))
    if P.alt_soundspeed:  # toolchain/mfc/case_validator.py:1394-1418
        if P.num_fluids != 2:
            raise UnsupportedCase("alt_soundspeed requires exactly 2 fluid components (the Kapila K coefficient is a two-fluid closure)")
        if P.avg_state != 2:
            raise UnsupportedCase("alt_soundspeed requires avg_state = 2")
        if P.riemann_solver == 1 and not P.hll_u_interface:
            raise UnsupportedCase("alt_soundspeed with HLL method 1 (interface alphas) is outside the hot path; use hll_u_interface = T or HLLC")
    P.low_Mach = int(_val(case, "low_Mach", 0) or 0)
    if P.low_Mach not in (0, 1, 2):
        raise UnsupportedCase("low_Mach must be 0, 1 or 2")
    if P.low_Mach == 2 and P.riemann_solver != 2:
        raise UnsupportedCase("low_Mach = 2 requires riemann_solver = 2 (HLLC)")  # m_checker_common / case_validator
    P.time_stepper = int(_val(case, "time_stepper", 3))
    P.dt = float(case["dt"])
    nd = P.num_dims
    for d, x in enumerate("xyz"):
        if d >= nd:
            continue
        P.bc[d][0] = int(case[f"bc_{x}%beg"])
        P.bc[d][1] = int(case[f"bc_{x}%end"])
        for a in range(3):
            P.bc_vb[d][a] = float(case.get(f"bc_{x}%vb{a + 1}", 0.0))
            P.bc_ve[d][a] = float(case.get(f"bc_{x}%ve{a + 1}", 0.0))
            P.vel_in[d][a] = float(case.get(f"bc_{x}%vel_in({a + 1})", 0.0))
            P.vel_out[d][a] = float(case.get(f"bc_{x}%vel_out({a + 1})", 0.0))
        P.grcbc_in[d] = _logical(case.get(f"bc_{x}%grcbc_in", "F"))
        P.grcbc_out[d] = _logical(case.get(f"bc_{x}%grcbc_out", "F"))
        P.grcbc_vel_out[d] = _logical(case.get(f"bc_{x}%grcbc_vel_out", "F"))
        P.pres_in[d] = float(case.get(f"bc_{x}%pres_in", 0.0))
        P.pres_out[d] = float(case.get(f"bc_{x}%pres_out", 0.0))
        for f in range(P.num_fluids):
            P.alpha_rho_in[d][f] = float(case.get(f"bc_{x}%alpha_rho_in({f + 1})", 0.0))
            P.alpha_in[d][f] = float(case.get(f"bc_{x}%alpha_in({f + 1})", 0.0))
    for f in range(P.num_fluids):
        P.fluid_gamma[f] = float(case[f"fluid_pp({f + 1})%gamma"])
        P.fluid_pi_inf[f] = float(case.get(f"fluid_pp({f + 1})%pi_inf", 0.0))
        P.fluid_qv[f] = float(case.get(f"fluid_pp({f + 1})%qv", 0.0))
    # viscous stresses (m_viscous.fpp; Re_size / Re_idx m_global_parameters.fpp:781-812): Newtonian fluids, Cartesian grids
    P.viscous = _logical(case.get("viscous", "F"))
    P.weno_Re_flux = _logical(case.get("weno_Re_flux", "F"))
    P.weno_avg = _logical(case.get("weno_avg", "F"))
    if P.viscous:
        if P.muscl_order > 0:
            raise UnsupportedCase("viscous with MUSCL reconstruction is outside the hot path")
        if P.riemann_solver == 5:
            raise UnsupportedCase("viscous with the Lax-Friedrichs solver (its own stress formulas, m_riemann_solver_lf.fpp:359-680) is outside the hot path")
        if P.weno_Re_flux:
            raise UnsupportedCase("weno_Re_flux = T (WENO-reconstructed velocity gradients) is outside the hot path; weno_Re_flux = F is available")
        if P.weno_order == 7:
            raise UnsupportedCase("viscous at weno_order = 7 is outside the hot path")
        for f in range(P.num_fluids):
            for i in range(2):
                key = f"fluid_pp({f + 1})%Re({i + 1})"
                P.fluid_Re[f][i] = float(case.get(key, -1.0e6))   # dflt_real
            if _logical(case.get(f"fluid_pp({f + 1})%non_newtonian", "F")):
                raise UnsupportedCase("non-Newtonian viscosity is outside the hot path")
        if not any(P.fluid_Re[f][i] > 0 for f in range(P.num_fluids) for i in range(2)):
            raise UnsupportedCase("viscous = T needs at least one fluid_pp(i)%Re(1) or %Re(2)")   # toolchain/mfc/case_validator.py
    ic = _val(case, "int_comp", 0)
    ic = (1 if _logical(ic) else 0) if isinstance(ic, str) and ic.strip().upper() in ("T", "F") else int(ic or 0)
    if ic not in (0, 1) and not (oracle_only and ic == 2):
        # MTHINC: in the CPU oracle and pinned by the reference's goldens (tests/test_oracle_next.py), not yet in the CUDA kernels
        raise UnsupportedCase("int_comp = 2 (MTHINC) is outside the hot path; 0 (off) and 1 (THINC) are available")
    if ic == 2 and P.n == 0:
        raise UnsupportedCase("int_comp = 2 (MTHINC) requires at least 2D (n > 0)")
    if ic >= 1:  # toolchain/mfc/case_validator.py:426-444
        if P.num_fluids != 2:
            raise UnsupportedCase("int_comp > 0 requires num_fluids = 2")
        if P.recon_order == 1:
            raise UnsupportedCase("int_comp must be 0 (off) when weno_order = 1 / muscl_order = 1")
    P.int_comp = ic
    P.ic_eps = float(case.get("ic_eps", 1e-4))
    P.ic_beta = float(case.get("ic_beta", 1.6))
    for d, x in enumerate("xyz"):  # body forces, m_body_forces.fpp:206-218 (a direction beyond num_dims is a checker error there)
        P.bf[d] = _logical(case.get(f"bf_{x}", "F"))
        if P.bf[d]:
            if d >= nd:
                raise UnsupportedCase(f"bf_{x} needs a {x} direction")
            P.bf_k[d], P.bf_w[d] = float(case.get(f"k_{x}", 0.0)), float(case.get(f"w_{x}", 0.0))
            P.bf_p[d], P.bf_g[d] = float(case.get(f"p_{x}", 0.0)), float(case.get(f"g_{x}", 0.0))
    if _logical(case.get("bf_spatial_support", "F")) or _logical(case.get("synthetic_turbulence", "F")):
        raise UnsupportedCase("bf_spatial_support / synthetic_turbulence are outside the hot path")
    P.device = device
    P.fast_math = fast_math
    check_params(P)
    return P


def check_params(P):
    """Run-time checks of the reference: src/simulation/m_checker.fpp:52-69 (cells per rank)."""
    need = 5 * P.recon_order   # num_stcls_min * weno_order (or muscl_order), order 1 included: m_checker.fpp:52-90
    if P.muscl_order > 0 and any(-12 <= P.bc[d][s] <= -5 for d in range(P.num_dims) for s in range(2)):
        raise UnsupportedCase("characteristic boundary conditions with MUSCL reconstruction are outside the hot path")
    if P.weno_order == 7 and P.muscl_order == 0:
        if any(-12 <= P.bc[d][s] <= -5 for d in range(P.num_dims) for s in range(2)):
            raise UnsupportedCase("characteristic boundary conditions at weno_order = 7 are outside the hot path (the reference's CBC tables stop at order 5)")
        if P.int_comp != 0 or P.null_weights:
            raise UnsupportedCase("int_comp / null_weights at weno_order = 7 are outside the hot path")
    for s, name in ((P.m, "m"), (P.n, "n"), (P.p, "p")):
        if s > 0 and s + 1 < need:
            raise UnsupportedCase(f"{name}+1 must be >= num_stcls_min*weno_order = {need}")
    if P.mapped_weno and P.wenoz:
        raise UnsupportedCase("mapped_weno and wenoz are mutually exclusive")
