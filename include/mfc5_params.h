/* mfc5_params.h -- the parameter contract of the 5-equation RHS hot path.
 *
 * One plain-C struct carrying every case.py / namelist parameter that reaches the hot path
 * (SURVEY.md Appendix C).  It is shared by the product library (include/mfc5.h, CUDA) and by the
 * CPU oracle (oracle/mfc5_oracle.h, test infrastructure).  Field names follow the reference's
 * namelist names (src/simulation/m_global_parameters.fpp:330-420, m_start_up.fpp:81-130).
 *
 * Variable ordering (reference m_global_parameters_common.fpp:125-141), 0-based here:
 *   [0, NF)            partial densities  alpha_k rho_k       (cont)
 *   [NF, NF+ND)        momenta (cons) / velocities (prim)     (mom)
 *   E = NF+ND          total energy (cons) / pressure (prim)
 *   (E, E+NF]          volume fractions alpha_k               (adv)
 *   sys_size N = 2*NF + ND + 1
 */
#ifndef MFC5_PARAMS_H
#define MFC5_PARAMS_H

#ifdef __cplusplus
extern "C" {
#endif

#define MFC5_MAX_FLUIDS 4
#define MFC5_MAX_EQ (2 * MFC5_MAX_FLUIDS + 4)

/* Boundary-condition codes, reference src/common/m_constants.fpp:97-113.  A value >= 0 in bc[][]
 * means "processor boundary, value = neighbour rank" (m_mpi_common.fpp:1496-1509). */
enum {
    MFC5_BC_PERIODIC = -1,
    MFC5_BC_REFLECTIVE = -2,
    MFC5_BC_GHOST_EXTRAP = -3,
    MFC5_BC_RIEMANN_EXTRAP = -4,
    MFC5_BC_CHAR_SLIP_WALL = -5,
    MFC5_BC_CHAR_NR_SUB_BUFFER = -6,
    MFC5_BC_CHAR_NR_SUB_INFLOW = -7,
    MFC5_BC_CHAR_NR_SUB_OUTFLOW = -8,
    MFC5_BC_CHAR_FF_SUB_OUTFLOW = -9,
    MFC5_BC_CHAR_CP_SUB_OUTFLOW = -10,
    MFC5_BC_CHAR_SUP_INFLOW = -11,
    MFC5_BC_CHAR_SUP_OUTFLOW = -12,
    MFC5_BC_SLIP_WALL = -15,
    MFC5_BC_NO_SLIP_WALL = -16,
    MFC5_BC_DIRICHLET = -17
};

typedef struct mfc5_params {
    /* grid extents of THIS rank: cells-1 per direction, n = 0 / p = 0 collapse dimensions */
    int m, n, p;
    int num_fluids;
    int model_eqns;      /* must be 2 (5-equation) */
    /* reconstruction */
    int weno_order;      /* 1, 3 or 5 */
    int mapped_weno;     /* WENO-M (Henrick) */
    int wenoz;           /* WENO-Z (Borges) ; neither => WENO-JS */
    int null_weights;
    double weno_eps;
    /* Riemann solver */
    int riemann_solver;  /* 1 = HLL, 2 = HLLC, 5 = Lax-Friedrichs (m_riemann_solver_lf.fpp) */
    int hll_u_interface; /* HLL alpha transport "method 2" */
    int avg_state;       /* 1 = Roe, 2 = arithmetic */
    int wave_speeds;     /* 1 = direct, 2 = pressure based */
    int mpp_lim;
    int mixture_err;
    /* time stepping */
    int time_stepper;    /* 1, 2, 3 (TVD RK order) */
    double dt;
    /* boundary conditions: bc[dir][0] = %beg, bc[dir][1] = %end */
    int bc[3][2];
    double bc_vb[3][3];  /* bc_<dir>%vb1..3 : wall velocity on the beg side */
    double bc_ve[3][3];  /* bc_<dir>%ve1..3 */
    /* GRCBC inputs (m_cbc.fpp:300-321) */
    int grcbc_in[3], grcbc_out[3], grcbc_vel_out[3];
    double vel_in[3][3], vel_out[3][3];
    double pres_in[3], pres_out[3];
    double alpha_rho_in[3][MFC5_MAX_FLUIDS], alpha_in[3][MFC5_MAX_FLUIDS];
    /* stiffened-gas EOS in the reference's form: gamma = 1/(g-1), pi_inf = g*pinf/(g-1) */
    double fluid_gamma[MFC5_MAX_FLUIDS];
    double fluid_pi_inf[MFC5_MAX_FLUIDS];
    double fluid_qv[MFC5_MAX_FLUIDS];
    /* product-side knobs (ignored by the oracle) */
    int device;          /* CUDA device ordinal */
    int fast_math;       /* 0: reference operation order; 1: division-light WENO weights */
    /* Riemann-solver extras; they occupy former reserved words, so the struct size is unchanged */
    int low_Mach;        /* 0 off; 1 pressure correction (HLL, HLLC); 2 velocity correction (HLLC): inline_riemann.fpp:71-93 */
    int alt_soundspeed;  /* Wood sound speed + K div(u) term in the alpha equations (2 fluids; HLLC or HLL with hll_u_interface) */
    /* MUSCL reconstruction (recon_type = 2, src/simulation/m_muscl.fpp): muscl_order 0 = off (WENO), 1, 2;
     * muscl_lim 0 unlimited, 1 minmod, 2 MC, 3 Van Albada, 4 Van Leer, 5 SUPERBEE; buff_size = muscl_order + 2 */
    int muscl_order, muscl_lim;
    double muscl_eps;
    /* TENO weights for weno_order = 5 (Fu et al. 2016; m_weno.fpp:1167-1187): cut-off teno_CT */
    int teno;
    int mp_weno;         /* monotonicity-preserving bounds on the WENO5 edge values (m_weno.fpp:1495-1601); reference-order kernels */
    double teno_CT;
    /* body forces (m_body_forces.fpp:206-218,294-385; applied after every RK stage update, m_time_steppers.fpp:533,713-737):
     * acceleration of direction d = bf_g[d] + bf_k[d] * sin(bf_w[d] * t - bf_p[d]) when bf[d] is set */
    int bf[3], reserved3;
    double bf_k[3], bf_w[3], bf_p[3], bf_g[3];
    /* interface compression: 0 off, 1 THINC (m_thinc.fpp:295-406, applied to the reconstructed alpha / alpha-rho of the two
     * fluids after WENO or MUSCL); ic_eps / ic_beta default to 1e-4 / 1.6 (m_constants.fpp:53-54) */
    int int_comp, reserved4;
    double ic_eps, ic_beta;
    /* WENO-Z power of order 7, alpha = d (1 + (tau / beta)^wenoz_q) (m_weno.fpp:1381-1389) */
    double wenoz_q;
    /* viscous stresses (src/simulation/m_viscous.fpp, m_riemann_state.fpp:805-1031): fluid_Re[k][0] = fluid_pp(k)%Re(1)
     * (shear), [k][1] = %Re(2) (bulk); a value <= 0 means "not set" (the reference's dflt_real).  weno_Re_flux = 0: velocity gradients
     * at the cell boundaries by finite differences (m_viscous.fpp:515-830); 1: WENO-reconstructed -- rejected by mfc5_create.  With viscous the
     * ghost region is 2 weno_polyn + 2 cells wide (m_helper_basic.fpp:127-132). */
    int viscous, weno_Re_flux, weno_avg, reserved5;
    double fluid_Re[MFC5_MAX_FLUIDS][2];
} mfc5_params;

#ifdef __cplusplus
}
#endif
#endif
