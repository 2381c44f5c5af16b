/* mfc5_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see mfc5_oracle.h).
 *
 * Unfused, array-for-array restatement of the reference's arithmetic, in the reference's
 * operation order.  Every routine cites the reference lines it follows (paths relative to
 * MFlowCode/MFC @ 0b234b19).  Build: oracle/Makefile (-O2 -ffp-contract=off so results do not
 * depend on FMA contraction).
 */
#include "mfc5_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SGM_EPS 1.e-16 /* src/common/m_constants.fpp:16 */
#define MAXF MFC5_MAX_FLUIDS
#define MAXN MFC5_MAX_EQ

struct orc_state {
    double mytime; /* simulation time of the state (body forces) */
    mfc5_params P;
    int nd, nf, N, b, polyn, ns; /* ns = weno_num_stencils + 1 */
    int s[3];                    /* m, n, p */
    int off[3], ext[3];
    long stride[3], fsz;
    int iE, iadv, imom;
    /* grid: scb[d][i], i in [-1-b, s+b]; ds[d][i], i in [-b, s+b] (pointers pre-offset) */
    double *scb_store[3], *ds_store[3];
    double *scb[3], *ds[3];
    /* WENO tables, index j in [-b+polyn, s+b-polyn] -> slot j + b (m_weno.fpp:170-893) */
    double *cL[3], *cR[3], *dL[3], *dR[3], *bt[3];
    int uniform[3];
    /* CBC coefficients per (dir, side) (m_cbc.fpp:327-408) */
    double fd[3][2][8], pi[3][2][2][3], cds[3][2][8], Del[3];
    /* fields */
    double *q1, *q2, *prim, *rhs, *vL, *vR, *flux, *fsrc;
    /* Dirichlet (-17) face values, bc_buffers of m_boundary_common.fpp:46-65: [v][t2][t1] over the transverse interior */
    double* dirich[3][2];
    /* MTHINC (int_comp = 2): unit normal and interface position per cell, m_thinc.fpp:36-38 */
    double *mth_n[3], *mth_d;
    /* viscous stresses (m_viscous.fpp, m_riemann_state.fpp:805-1031): cell centres scc[d][i], i in [-b, s+b]; velocity-gradient
     * fields at the cell boundaries dqL / dqR[derivative direction][face-normal direction] (nd velocity components each, the
     * reference's dqL_prim_d<x|y|z>_n(normal)%vf(mom)); interface Reynolds numbers and velocities exported by the Riemann solver
     * (Re_avg_rsx_vf, vel_src_rsx_vf); viscous source flux of the momentum and energy equations (flux_src_n(dir)%vf(mom..E)) */
    double *scc_store[3], *scc[3];
    double *dqL[3][3], *dqR[3][3], *re_avg, *vel_src, *fvis;
    int re_size[2], re_idx[2][MAXF];
    double res[2][MAXF];
};

static inline long IDX(const orc_state* S, int i, int j, int k) {
    return (long)(i + S->off[0]) + (long)S->ext[0] * ((long)(j + S->off[1]) + (long)S->ext[1] * (long)(k + S->off[2]));
}
static inline double sq(double x) { return x * x; }
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* exp() for THINC, restated from the published fdlibm algorithm (Sun's e_exp.c: Cody-Waite reduction x = k ln2 + r,
 * |r| <= 0.5 ln2, exp(r) = 1 + 2r/(R(r^2) - r) with a degree-5 Remez polynomial; error < 1 ulp) instead of libm's, so that
 * the strict CUDA build (same operations, no FMA contraction) reproduces the oracle bit for bit: CUDA's exp and glibc's
 * differ in the last bit for some arguments, and THINC's thresholds amplify that.  Arguments here are |x| <= ~2 ic_beta. */
#ifndef ORC_LIBM_EXP
static double orc_exp(double x) {
    static const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00,
                        P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                        P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (!(fabs(x) < 700.0)) return exp(x); /* outside the range THINC can produce */
    int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
    double hi = x - (double)k * ln2hi, lo = (double)k * ln2lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    return ldexp(y, k);
}
#else
#define orc_exp exp
#endif

/* Monotonicity-preserving bounds on the WENO5 edge values (mp_weno), s_preserve_monotonicity m_weno.fpp:1495-1601.
 * mp_md(a, b): the reference's median/curvature expression on two neighbouring second differences. */
static inline double fsgn1(double b) { return copysign(1.0, b); }
static inline double mp_md(double a, double b) {
    double s = fsgn1(4.0 * a - b);
    double mn = dmin(dmin(dmin(fabs(4.0 * a - b), fabs(a)), fabs(4.0 * b - a)), fabs(b));
    return (s + fsgn1(4.0 * b - a)) * fabs((s + fsgn1(a)) * (s + fsgn1(b))) * mn / 8.0;
}
static void mp_limit(double vm2, double vm1, double vp0, double vp1, double vp2, double* vL, double* vR) {
    const double alpha_mp = 2.0, beta_mp = 4.0 / 3.0;
    double dm = vp0 + vm2 - vm1 * 2.0, d0 = vp1 + vm1 - vp0 * 2.0, dp = vp2 + vp0 - vp1 * 2.0;
    double d_MD = mp_md(dm, d0), d_LC = mp_md(d0, dp);
    double UL = vp0 - (vp1 - vp0) * alpha_mp;
    double MD = (vp0 + vm1 - d_MD) * 5.e-1;
    double LC = vp0 - (vp1 - vp0) * 5.e-1 + beta_mp * d_LC;
    double vmin = dmax(dmin(dmin(vp0, vm1), MD), dmin(dmin(vp0, UL), LC));
    double vmax = dmin(dmax(dmax(vp0, vm1), MD), dmax(dmax(vp0, UL), LC));
    *vL = *vL + (copysign(5.e-1, vmin - *vL) + copysign(5.e-1, vmax - *vL)) * dmin(fabs(vmin - *vL), fabs(vmax - *vL));
    d_MD = mp_md(d0, dp); d_LC = mp_md(dm, d0);
    UL = vp0 + (vp0 - vm1) * alpha_mp;
    MD = (vp0 + vp1 - d_MD) * 5.e-1;
    LC = vp0 + (vp0 - vm1) * 5.e-1 + beta_mp * d_LC;
    vmin = dmax(dmin(dmin(vp0, vp1), MD), dmin(dmin(vp0, UL), LC));
    vmax = dmin(dmax(dmax(vp0, vp1), MD), dmax(dmax(vp0, UL), LC));
    *vR = *vR + (copysign(5.e-1, vmin - *vR) + copysign(5.e-1, vmax - *vR)) * dmin(fabs(vmin - *vR), fabs(vmax - *vR));
}
/* Fortran sign(a, b): |a| with the sign of b (b = +0 -> +, b = -0 -> -) */
static inline double fsign(double a, double b) { return copysign(fabs(a), b); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

orc_state* orc_create(const mfc5_params* P) {
    orc_state* S = (orc_state*)calloc(1, sizeof(orc_state));
    S->P = *P;
    S->s[0] = P->m; S->s[1] = P->n; S->s[2] = P->p;
    S->nd = 1 + (P->n > 0) + (P->p > 0);
    S->nf = P->num_fluids;
    S->N = 2 * S->nf + S->nd + 1;
    if (P->muscl_order > 0) S->P.weno_order = 2 * P->muscl_order + 1; /* stencil / buffer geometry: buff_size = muscl_order + 2 */
    S->polyn = (S->P.weno_order - 1) / 2;
    S->ns = S->polyn + 1;
    /* buff_size = weno_polyn + 2 for inviscid WENO, 2 weno_polyn + 2 with viscous: src/common/m_helper_basic.fpp:127-135 */
    S->b = S->polyn + 2;
    if (P->viscous && P->muscl_order == 0) S->b = 2 * S->polyn + 2;
    for (int d = 0; d < 3; d++) {
        S->off[d] = (d < S->nd) ? S->b : 0;
        S->ext[d] = S->s[d] + 1 + 2 * S->off[d];
    }
    S->stride[0] = 1; S->stride[1] = S->ext[0]; S->stride[2] = (long)S->ext[0] * S->ext[1];
    S->fsz = (long)S->ext[0] * S->ext[1] * S->ext[2];
    S->imom = S->nf; S->iE = S->nf + S->nd; S->iadv = S->iE + 1;
    size_t nb = (size_t)S->N * S->fsz * sizeof(double);
    S->q1 = (double*)calloc(1, nb); S->q2 = (double*)calloc(1, nb); S->prim = (double*)calloc(1, nb);
    S->rhs = (double*)calloc(1, nb); S->vL = (double*)calloc(1, nb); S->vR = (double*)calloc(1, nb);
    S->flux = (double*)calloc(1, nb); S->fsrc = (double*)calloc(1, (size_t)MAXF * S->fsz * sizeof(double));
    if (P->viscous) {
        /* Re_size / Re_idx / Res: m_global_parameters.fpp:781-812, m_variables_conversion.fpp:315-322 */
        for (int i = 0; i < 2; i++) {
            S->re_size[i] = 0;
            for (int f = 0; f < S->nf; f++)
                if (P->fluid_Re[f][i] > 0) { S->re_idx[i][S->re_size[i]] = f; S->res[i][S->re_size[i]] = P->fluid_Re[f][i]; S->re_size[i]++; }
        }
        size_t nv = (size_t)S->nd * S->fsz * sizeof(double);
        for (int dd = 0; dd < S->nd; dd++)
            for (int n = 0; n < S->nd; n++) { S->dqL[dd][n] = (double*)calloc(1, nv); S->dqR[dd][n] = (double*)calloc(1, nv); }
        S->re_avg = (double*)calloc(1, 2 * S->fsz * sizeof(double));
        S->vel_src = (double*)calloc(1, 3 * S->fsz * sizeof(double));
        S->fvis = (double*)calloc(1, (size_t)(S->nd + 1) * S->fsz * sizeof(double));
    }
    return S;
}

void orc_destroy(orc_state* S) {
    if (!S) return;
    for (int d = 0; d < 3; d++) {
        free(S->scb_store[d]); free(S->ds_store[d]);
        free(S->cL[d]); free(S->cR[d]); free(S->dL[d]); free(S->dR[d]); free(S->bt[d]); free(S->mth_n[d]);
    }
    for (int d = 0; d < 3; d++) { free(S->dirich[d][0]); free(S->dirich[d][1]); free(S->scc_store[d]); }
    for (int dd = 0; dd < 3; dd++) for (int n = 0; n < 3; n++) { free(S->dqL[dd][n]); free(S->dqR[dd][n]); }
    free(S->re_avg); free(S->vel_src); free(S->fvis);
    free(S->q1); free(S->q2); free(S->prim); free(S->rhs); free(S->vL); free(S->vR); free(S->flux); free(S->fsrc); free(S->mth_d);
    free(S);
}
int orc_sys_size(const orc_state* S) { return S->N; }
int orc_buff_size(const orc_state* S) { return S->b; }

/* ---------------------------------------------------------------------------------------------
 * Grid ghost metrics: src/common/m_boundary_common.fpp:404-466 (physical BCs),
 * src/common/m_mpi_common.fpp:1745-1796 (processor boundaries); interior widths as
 * src/simulation/m_start_up.fpp:190-191.
 * ------------------------------------------------------------------------------------------- */
void orc_set_grid(orc_state* S, int d, const double* s_cb_in, const double* gbeg, const double* gend) {
    int s = S->s[d], b = S->b;
    S->scb_store[d] = (double*)calloc((size_t)(s + 2 + 2 * b), sizeof(double));
    S->ds_store[d] = (double*)calloc((size_t)(s + 1 + 2 * b), sizeof(double));
    double* cb = S->scb_store[d] + 1 + b; /* cb[-1-b .. s+b] */
    double* w = S->ds_store[d] + b;       /* w[-b .. s+b] */
    S->scb[d] = cb; S->ds[d] = w;
    for (int i = -1; i <= s; i++) cb[i] = s_cb_in[i + 1];
    for (int i = 0; i <= s; i++) w[i] = cb[i] - cb[i - 1];
    int bcb = S->P.bc[d][0], bce = S->P.bc[d][1];
    for (int i = 1; i <= b; i++) {
        if (bcb >= 0) w[-i] = gbeg[b - i]; /* neighbour's last b widths, increasing index */
        else {
            int src = (bcb == MFC5_BC_PERIODIC) ? s - i + 1 : (bcb == MFC5_BC_REFLECTIVE) ? i - 1 : 0;
            w[-i] = w[src];
        }
    }
    for (int i = 1; i <= b; i++) {
        if (bce >= 0) w[s + i] = gend[i - 1];
        else {
            int src = (bce == MFC5_BC_PERIODIC) ? i - 1 : (bce == MFC5_BC_REFLECTIVE) ? s - i + 1 : s;
            w[s + i] = w[src];
        }
    }
    for (int i = 1; i <= b; i++) cb[-1 - i] = cb[-i] - w[-i];
    for (int i = 1; i <= b; i++) cb[s + i] = cb[s + i - 1] + w[s + i];
    /* cell centres: interior m_start_up.fpp:191, ghosts by the recurrence of s_populate_grid_bc_direction
     * (m_boundary_common.fpp:449-465) */
    S->scc_store[d] = (double*)calloc((size_t)(s + 1 + 2 * b), sizeof(double));
    double* cc = S->scc_store[d] + b;
    S->scc[d] = cc;
    for (int i = 0; i <= s; i++) cc[i] = cb[i - 1] + w[i] / 2.0;
    for (int i = 1; i <= b; i++) cc[-i] = cc[1 - i] - (w[1 - i] + w[-i]) / 2.0;
    for (int i = 1; i <= b; i++) cc[s + i] = cc[s + i - 1] + (w[s + i - 1] + w[s + i]) / 2.0;
}

/* ---------------------------------------------------------------------------------------------
 * WENO coefficient tables: src/simulation/m_weno.fpp:196-215 (WENO3), :233-341 (WENO5),
 * null_weights :219-229,345-361, uniform-grid detection :869-879.
 * ------------------------------------------------------------------------------------------- */
#define GK_MAX 4
static void g_weno_tables(int k, const double* xb, double cLo[][GK_MAX - 1], double cRo[][GK_MAX - 1], double* dLo, double* dRo,
                          double bto[][6]);
static void weno_coefficients(orc_state* S, int d) {
    int s = S->s[d], b = S->b, pn = S->polyn, ns = S->ns;
    int nslot = s + 1 + 2 * b;
    S->cL[d] = (double*)calloc((size_t)nslot * ns * (pn > 0 ? pn : 1), sizeof(double));
    S->cR[d] = (double*)calloc((size_t)nslot * ns * (pn > 0 ? pn : 1), sizeof(double));
    S->dL[d] = (double*)calloc((size_t)nslot * ns, sizeof(double));
    S->dR[d] = (double*)calloc((size_t)nslot * ns, sizeof(double));
    S->bt[d] = (double*)calloc((size_t)nslot * ns * (pn == 3 ? 6 : 3), sizeof(double));
    const double* scb = S->scb[d];
#define C(k) scb[i + (k)]
#define CL(j, q, r) S->cL[d][(((j) + b) * ns + (q)) * pn + (r)]
#define CR(j, q, r) S->cR[d][(((j) + b) * ns + (q)) * pn + (r)]
#define DL(q, j) S->dL[d][((j) + b) * ns + (q)]
#define DR(q, j) S->dR[d][((j) + b) * ns + (q)]
#define BT(j, q, r) S->bt[d][(((j) + b) * ns + (q)) * 3 + (r)]
    int bcb = S->P.bc[d][0], bce = S->P.bc[d][1];
    if (S->P.weno_order == 3) {
        for (int i = -b - 1 + pn; i <= s + b - 1 - pn; i++) {
            int j = i + 1;
            CR(j, 0, 0) = (C(0) - C(1)) / (C(0) - C(2));
            CR(j, 1, 0) = (C(0) - C(1)) / (C(-1) - C(1));
            CL(j, 0, 0) = -CR(j, 0, 0);
            CL(j, 1, 0) = -CR(j, 1, 0);
            DR(0, j) = (C(-1) - C(1)) / (C(-1) - C(2));
            DL(0, j) = (C(-1) - C(0)) / (C(-1) - C(2));
            DR(1, j) = 1.0 - DR(0, j);
            DL(1, j) = 1.0 - DL(0, j);
            BT(j, 0, 0) = 4.0 * sq(C(0) - C(1)) / sq(C(0) - C(2));
            BT(j, 1, 0) = 4.0 * sq(C(0) - C(1)) / sq(C(-1) - C(1));
        }
        if (S->P.null_weights) {
            if (bcb == MFC5_BC_RIEMANN_EXTRAP) { DR(1, 0) = 0; DR(0, 0) = 1; DL(1, 0) = 0; DL(0, 0) = 1; }
            if (bce == MFC5_BC_RIEMANN_EXTRAP) { DR(0, s) = 0; DR(1, s) = 1; DL(0, s) = 0; DL(1, s) = 1; }
        }
    } else if (S->P.weno_order == 5) {
        for (int i = -b - 1 + pn; i <= s + b - 1 - pn; i++) {
            int j = i + 1;
            CR(j, 0, 0) = ((C(0) - C(1)) * (C(1) - C(2))) / ((C(0) - C(3)) * (C(3) - C(1)));
            CR(j, 1, 0) = ((C(-1) - C(1)) * (C(1) - C(0))) / ((C(-1) - C(2)) * (C(2) - C(0)));
            CR(j, 1, 1) = ((C(0) - C(1)) * (C(1) - C(2))) / ((C(-1) - C(1)) * (C(-1) - C(2)));
            CR(j, 2, 1) = ((C(0) - C(1)) * (C(1) - C(-1))) / ((C(-2) - C(0)) * (C(-2) - C(1)));
            CL(j, 0, 0) = ((C(1) - C(0)) * (C(0) - C(2))) / ((C(0) - C(3)) * (C(3) - C(1)));
            CL(j, 1, 0) = ((C(0) - C(-1)) * (C(0) - C(1))) / ((C(-1) - C(2)) * (C(0) - C(2)));
            CL(j, 1, 1) = ((C(1) - C(0)) * (C(0) - C(2))) / ((C(-1) - C(1)) * (C(-1) - C(2)));
            CL(j, 2, 1) = ((C(-1) - C(0)) * (C(0) - C(1))) / ((C(-2) - C(0)) * (C(-2) - C(1)));

            CR(j, 0, 1) = ((C(0) - C(2)) + (C(1) - C(3))) / ((C(0) - C(2)) * (C(0) - C(3))) * ((C(0) - C(1)));
            CR(j, 2, 0) = ((C(-2) - C(1)) + (C(-1) - C(1))) / ((C(-1) - C(1)) * (C(1) - C(-2))) * ((C(1) - C(0)));
            CL(j, 0, 1) = ((C(0) - C(2)) + (C(0) - C(3))) / ((C(0) - C(2)) * (C(0) - C(3))) * ((C(1) - C(0)));
            CL(j, 2, 0) = ((C(-2) - C(0)) + (C(-1) - C(1))) / ((C(-2) - C(1)) * (C(1) - C(-1))) * ((C(0) - C(1)));

            DR(0, j) = ((C(-2) - C(1)) * (C(1) - C(-1))) / ((C(-2) - C(3)) * (C(3) - C(-1)));
            DR(2, j) = ((C(1) - C(2)) * (C(1) - C(3))) / ((C(-2) - C(2)) * (C(-2) - C(3)));
            DL(0, j) = ((C(-2) - C(0)) * (C(0) - C(-1))) / ((C(-2) - C(3)) * (C(3) - C(-1)));
            DL(2, j) = ((C(0) - C(2)) * (C(0) - C(3))) / ((C(-2) - C(2)) * (C(-2) - C(3)));
            DR(1, j) = 1.0 - DR(0, j) - DR(2, j);
            DL(1, j) = 1.0 - DL(0, j) - DL(2, j);

            BT(j, 0, 0) = 4.0 * sq(C(0) - C(1)) * (10.0 * sq(C(1) - C(0)) + (C(1) - C(0)) * (C(2) - C(1)) + sq(C(2) - C(1))) /
                          (sq(C(0) - C(3)) * sq(C(1) - C(3)));
            BT(j, 0, 1) = 4.0 * sq(C(0) - C(1)) *
                          (19.0 * sq(C(1) - C(0)) - (C(1) - C(0)) * (C(3) - C(1)) +
                           2.0 * (C(2) - C(0)) * ((C(2) - C(0)) + (C(3) - C(1)))) /
                          ((C(0) - C(2)) * sq(C(0) - C(3)) * (C(3) - C(1)));
            BT(j, 0, 2) = 4.0 * sq(C(0) - C(1)) *
                          (10.0 * sq(C(1) - C(0)) + (C(1) - C(0)) * ((C(2) - C(0)) + (C(3) - C(1))) +
                           sq((C(2) - C(0)) + (C(3) - C(1)))) /
                          (sq(C(0) - C(2)) * sq(C(0) - C(3)));
            BT(j, 1, 0) = 4.0 * sq(C(0) - C(1)) * (10.0 * sq(C(1) - C(0)) + sq(C(0) - C(-1)) + (C(0) - C(-1)) * (C(1) - C(0))) /
                          (sq(C(-1) - C(2)) * sq(C(0) - C(2)));
            BT(j, 1, 1) = 4.0 * sq(C(0) - C(1)) *
                          ((C(0) - C(1)) * ((C(0) - C(-1)) + 20.0 * (C(1) - C(0))) +
                           (2.0 * (C(0) - C(-1)) + (C(1) - C(0))) * (C(2) - C(0))) /
                          ((C(1) - C(-1)) * sq(C(-1) - C(2)) * (C(2) - C(0)));
            BT(j, 1, 2) = 4.0 * sq(C(0) - C(1)) * (10.0 * sq(C(1) - C(0)) + (C(1) - C(0)) * (C(2) - C(1)) + sq(C(2) - C(1))) /
                          (sq(C(-1) - C(1)) * sq(C(-1) - C(2)));
            BT(j, 2, 0) = 4.0 * sq(C(0) - C(1)) *
                          (12.0 * sq(C(1) - C(0)) + sq((C(0) - C(-2)) + (C(0) - C(-1))) +
                           3.0 * ((C(0) - C(-2)) + (C(0) - C(-1))) * (C(1) - C(0))) /
                          (sq(C(-2) - C(1)) * sq(C(-1) - C(1)));
            BT(j, 2, 1) = 4.0 * sq(C(0) - C(1)) *
                          (19.0 * sq(C(1) - C(0)) + ((C(0) - C(-2)) * (C(0) - C(1))) +
                           2.0 * (C(1) - C(-1)) * ((C(0) - C(-2)) + (C(1) - C(-1)))) /
                          ((C(-2) - C(0)) * sq(C(-2) - C(1)) * (C(1) - C(-1)));
            BT(j, 2, 2) = 4.0 * sq(C(0) - C(1)) * (10.0 * sq(C(1) - C(0)) + sq(C(0) - C(-1)) + (C(0) - C(-1)) * (C(1) - C(0))) /
                          (sq(C(-2) - C(0)) * sq(C(-2) - C(1)));
        }
        if (S->P.null_weights) {
            if (bcb == MFC5_BC_RIEMANN_EXTRAP) {
                DR(1, 0) = 0; DR(2, 0) = 0; DR(0, 0) = 1; DL(1, 0) = 0; DL(2, 0) = 0; DL(0, 0) = 1;
                DR(2, 1) = 0; { double t = DR(0, 1) + DR(1, 1) + DR(2, 1); for (int q = 0; q < 3; q++) DR(q, 1) /= t; }
                DL(2, 1) = 0; { double t = DL(0, 1) + DL(1, 1) + DL(2, 1); for (int q = 0; q < 3; q++) DL(q, 1) /= t; }
            }
            if (bce == MFC5_BC_RIEMANN_EXTRAP) {
                DR(0, s - 1) = 0; { double t = DR(0, s - 1) + DR(1, s - 1) + DR(2, s - 1); for (int q = 0; q < 3; q++) DR(q, s - 1) /= t; }
                DL(0, s - 1) = 0; { double t = DL(0, s - 1) + DL(1, s - 1) + DL(2, s - 1); for (int q = 0; q < 3; q++) DL(q, s - 1) /= t; }
                DR(0, s) = 0; DR(1, s) = 0; DR(2, s) = 1; DL(0, s) = 0; DL(1, s) = 0; DL(2, s) = 1;
            }
        }
    }
    if (S->P.weno_order == 7) { /* m_weno.fpp:362-851, by the general construction (g_weno_tables, defined below) */
        for (int j = -b + pn; j <= s + b - pn; j++) {
            double xb[8], cl[GK_MAX][GK_MAX - 1], cr[GK_MAX][GK_MAX - 1], dl[GK_MAX], dr[GK_MAX], bt[GK_MAX][6];
            for (int m = 0; m < 8; m++) xb[m] = scb[j - 4 + m];
            g_weno_tables(4, xb, cl, cr, dl, dr, bt);
            for (int q = 0; q < 4; q++) {
                for (int r = 0; r < 3; r++) { CL(j, q, r) = cl[q][r]; CR(j, q, r) = cr[q][r]; }
                /* the reference numbers the ideal weights, polynomials and betas of order 7 from the LEFTMOST stencil (d_cb(0) =
                 * 1/35 on a uniform grid) while poly_coef / beta_coef keep the rightmost-first index: d(q) belongs to sidx 3 - q */
                DL(q, j) = dl[3 - q];
                DR(q, j) = dr[3 - q];
                for (int r = 0; r < 6; r++) S->bt[d][(((j) + b) * ns + q) * 6 + r] = bt[q][r];
            }
        }
    }
    /* uniform-grid detection, m_weno.fpp:869-879 (epsilon(h0) = 2^-52) */
    S->uniform[d] = 1;
    if (s > 0) {
        double h0 = (scb[s] - scb[0]) / (double)s;
        for (int i = 0; i <= s - 1; i++)
            if (fabs((scb[i + 1] - scb[i]) - h0) > sqrt(2.220446049250313e-16) * fabs(h0)) { S->uniform[d] = 0; break; }
    }
#undef C
}


/* ---------------------------------------------------------------------------------------------
 * WENO tables by the general construction, used for weno_order = 7 (m_weno.fpp:362-851 lists the same quantities as ~500
 * lines of closed forms; restated here from the published definitions they come from):
 *   - candidate polynomials: Shu (1997), eq. 2.20 -- the degree-(k-1) polynomial whose cell averages match k cells is the
 *     derivative of the Lagrange interpolant of the primitive function through the k+1 cell boundaries;
 *   - written, like the reference, on the cell-value differences dvd(t) = v(i+t+1) - v(i+t):  p = v(i) + sum_t a_t dvd(t);
 *   - ideal weights: the combination of the k candidates that reproduces the (2k-1)-cell polynomial at the cell boundary;
 *   - smoothness indicators: Jiang & Shu / Shu eq. 2.61, beta = sum_{l=1}^{k-1} dx_i^(2l-1) int_cell (d^l p / dx^l)^2 dx, as
 *     the quadratic form in the stencil's dvd's (beta_coef: squares and doubled cross terms, highest dvd first).
 * Checked against the transcribed WENO5 closed forms on a stretched grid (orc_selftest_weno_generic: 1e-13) and pinned by
 * the reference's WENO7 goldens.  Stencil index sidx = 0 is the rightmost one (cells i .. i+k-1), as in poly_coef_cb.
 * ------------------------------------------------------------------------------------------- */
typedef struct { double c[8]; int n; } gpoly; /* c[0] + c[1] x + ..., n = number of coefficients */
static gpoly gp_mul_lin(gpoly a, double r, double scale) { /* a * (x - r) * scale */
    gpoly o; o.n = a.n + 1;
    for (int i = 0; i < o.n; i++) o.c[i] = 0.0;
    for (int i = 0; i < a.n; i++) { o.c[i + 1] += a.c[i] * scale; o.c[i] -= a.c[i] * r * scale; }
    return o;
}
static gpoly gp_deriv(gpoly a) {
    gpoly o; o.n = a.n > 1 ? a.n - 1 : 1; o.c[0] = 0.0;
    for (int i = 1; i < a.n; i++) o.c[i - 1] = a.c[i] * (double)i;
    return o;
}
static gpoly gp_axpy(gpoly y, double a, gpoly x) { /* y + a x */
    gpoly o; o.n = y.n > x.n ? y.n : x.n;
    for (int i = 0; i < o.n; i++) o.c[i] = (i < y.n ? y.c[i] : 0.0) + a * (i < x.n ? x.c[i] : 0.0);
    return o;
}
static double gp_eval(gpoly a, double x) { double v = 0.0; for (int i = a.n - 1; i >= 0; i--) v = v * x + a.c[i]; return v; }
static double gp_int_prod(gpoly a, gpoly b, double L) { /* int_0^L a b dx */
    double prod[16]; int n = a.n + b.n - 1;
    for (int i = 0; i < n; i++) prod[i] = 0.0;
    for (int i = 0; i < a.n; i++) for (int j = 0; j < b.n; j++) prod[i + j] += a.c[i] * b.c[j];
    double v = 0.0, Lp = L;
    for (int i = 0; i < n; i++) { v += prod[i] * Lp / (double)(i + 1); Lp *= L; }
    return v;
}
/* c_j(x), j = 0..kk-1: weights of the cell averages in the reconstruction polynomial over the cells bounded by X[0..kk] */
static void g_cell_weights(const double* X, int kk, gpoly* cj) {
    gpoly dL[2 * GK_MAX]; /* L_m'(x), m = 0..kk */
    for (int m = 0; m <= kk; m++) {
        gpoly l; l.n = 1; l.c[0] = 1.0;
        for (int n = 0; n <= kk; n++) if (n != m) l = gp_mul_lin(l, X[n], 1.0 / (X[m] - X[n]));
        dL[m] = gp_deriv(l);
    }
    for (int j = 0; j < kk; j++) {
        gpoly acc; acc.n = 1; acc.c[0] = 0.0;
        for (int m = j + 1; m <= kk; m++) acc = gp_axpy(acc, 1.0, dL[m]);
        for (int i = 0; i < acc.n; i++) acc.c[i] *= (X[j + 1] - X[j]);
        cj[j] = acc;
    }
}
/* tables of cell i for order 2k-1; xb = boundaries x_{i-k+1-1/2} .. x_{i+k-1+1/2} (2k values) relative to any origin.
 * out: cL/cR[sidx][r] (r = 0: highest dvd), dL/dR[sidx], bt[sidx][0..k(k-1)/2-1] */
static void g_weno_tables(int k, const double* xb, double cLo[][GK_MAX - 1], double cRo[][GK_MAX - 1], double* dLo, double* dRo,
                          double bto[][6]) {
    double X[2 * GK_MAX];
    const double x0 = xb[k - 1];              /* left boundary of cell i */
    for (int m = 0; m < 2 * k; m++) X[m] = xb[m] - x0;
    const double dxi = X[k] - X[k - 1];       /* = X[k] */
    gpoly cw[GK_MAX][GK_MAX];                 /* cw[sidx][local cell] */
    for (int sidx = 0; sidx < k; sidx++) {
        const int l = sidx;                   /* cell i is local cell l of the stencil starting at global cell i - l ...      */
        (void)l;
    }
    /* stencil with cell i at local position l covers boundaries X[k-1-l .. 2k-1-l]; sidx = 0 <-> l = 0 (rightmost) */
    for (int l = 0; l < k; l++) {
        g_cell_weights(X + (k - 1 - l), k, cw[l]);
        gpoly a[GK_MAX - 1];                  /* a[e]: coefficient polynomial of the local difference v(e+1) - v(e) */
        for (int e = 0; e < k - 1; e++) {
            gpoly acc; acc.n = 1; acc.c[0] = 0.0;
            if (e >= l) { for (int j = e + 1; j < k; j++) acc = gp_axpy(acc, 1.0, cw[l][j]); }
            else        { for (int j = 0; j <= e; j++) acc = gp_axpy(acc, -1.0, cw[l][j]); }
            a[e] = acc;
        }
        for (int r = 0; r < k - 1; r++) {     /* r = 0 <-> highest difference e = k-2 */
            const int e = k - 2 - r;
            cLo[l][r] = gp_eval(a[e], 0.0);
            cRo[l][r] = gp_eval(a[e], dxi);
        }
        /* beta quadratic form */
        double B[GK_MAX - 1][GK_MAX - 1];
        for (int e = 0; e < k - 1; e++) for (int f = 0; f < k - 1; f++) B[e][f] = 0.0;
        gpoly da[GK_MAX - 1];
        for (int e = 0; e < k - 1; e++) da[e] = a[e];
        double w = dxi;                       /* dx^(2l-1) */
        for (int ord = 1; ord <= k - 1; ord++) {
            for (int e = 0; e < k - 1; e++) da[e] = gp_deriv(da[e]);
            for (int e = 0; e < k - 1; e++) for (int f = 0; f < k - 1; f++) B[e][f] += w * gp_int_prod(da[e], da[f], dxi);
            w *= dxi * dxi;
        }
        int n = 0;
        for (int r = 0; r < k - 1; r++) for (int r2 = r; r2 < k - 1; r2++) {
            const int e = k - 2 - r, f = k - 2 - r2;
            bto[l][n++] = (r == r2) ? B[e][e] : 2.0 * B[e][f];
        }
    }
    /* ideal weights from the (2k-1)-cell polynomial */
    gpoly cbig[2 * GK_MAX];
    g_cell_weights(X, 2 * k - 1, cbig);
    for (int side = 0; side < 2; side++) {
        const double xe = side == 0 ? 0.0 : dxi;
        double* dd = side == 0 ? dLo : dRo;
        /* global cell g = 0 .. 2k-2 (g = k-1 is cell i); stencil l covers g = k-1-l .. 2k-2-l */
        for (int l = k - 1; l >= 0; l--) {
            const int g = k - 1 - l;          /* the leftmost cell of stencil l; stencils l' > l also contain it */
            double rhs = gp_eval(cbig[g], xe);
            for (int l2 = k - 1; l2 > l; l2--) rhs -= dd[l2] * gp_eval(cw[l2][g - (k - 1 - l2)], xe);
            dd[l] = rhs / gp_eval(cw[l][0], xe);
        }
    }
}

/* self-test: the general construction for k = 3 against the transcribed WENO5 closed forms of direction d; returns the largest
 * deviation relative to the largest table entry of each kind */
double orc_selftest_weno_generic(orc_state* S, int d) {
    if (S->P.weno_order != 5) return -1.0;
    const int s = S->s[d], b = S->b, pn = S->polyn, ns = S->ns;
    double worst = 0.0;
    for (int j = -b + pn; j <= s + b - pn; j++) {
        double xb[6], cl[GK_MAX][GK_MAX - 1], cr[GK_MAX][GK_MAX - 1], dl[GK_MAX], dr[GK_MAX], bt[GK_MAX][6];
        for (int m = 0; m < 6; m++) xb[m] = S->scb[d][j - 3 + m];
        g_weno_tables(3, xb, cl, cr, dl, dr, bt);
        for (int q = 0; q < 3; q++) {
            for (int r = 0; r < 2; r++) {
                worst = fmax(worst, fabs(cl[q][r] - S->cL[d][((j + b) * ns + q) * pn + r]));
                worst = fmax(worst, fabs(cr[q][r] - S->cR[d][((j + b) * ns + q) * pn + r]));
            }
            worst = fmax(worst, fabs(dl[q] - S->dL[d][(j + b) * ns + q]));
            worst = fmax(worst, fabs(dr[q] - S->dR[d][(j + b) * ns + q]));
            for (int r = 0; r < 3; r++) worst = fmax(worst, fabs(bt[q][r] - S->bt[d][((j + b) * ns + q) * 3 + r]) / 10.0);
        }
    }
    return worst;
}

/* CBC coefficients: src/simulation/m_cbc.fpp:327-408 (fd_coef, pi_coef), :410-461 (ds), Del :312-313 */
static void cbc_coefficients(orc_state* S, int d, int side) {
    int s = S->s[d], b = S->b;
    double* ds = S->cds[d][side];
    for (int i = 0; i <= b; i++) ds[i] = side == 0 ? S->ds[d][i] : S->ds[d][s - i];
    double scb[10];
    scb[0] = 0.0;
    for (int i = 0; i <= b; i++) scb[i + 1] = scb[i] + ds[i];
    double* fd = S->fd[d][side];
    for (int i = 0; i < 8; i++) fd[i] = 0.0;
    if (S->P.weno_order == 1) {
        fd[0] = -2.0 / (ds[0] + ds[1]);
        fd[1] = -fd[0];
    } else if (S->P.weno_order == 3) {
        fd[0] = -6.0 / (3.0 * ds[0] + 2.0 * ds[1] - ds[2]);
        fd[1] = -4.0 * fd[0] / 3.0;
        fd[2] = fd[0] / 3.0;
        S->pi[d][side][0][0] = (scb[0] - scb[1]) / (scb[0] - scb[2]);
    } else {
        fd[0] = -50.0 / (25.0 * ds[0] + 2.0 * ds[1] - 1.e1 * ds[2] + 1.e1 * ds[3] - 3.0 * ds[4]);
        fd[1] = -48.0 * fd[0] / 25.0;
        fd[2] = 36.0 * fd[0] / 25.0;
        fd[3] = -16.0 * fd[0] / 25.0;
        fd[4] = 3.0 * fd[0] / 25.0;
        double(*pi)[3] = S->pi[d][side];
        pi[0][0] = ((scb[0] - scb[1]) * (scb[1] - scb[2]) * (scb[1] - scb[3])) /
                   ((scb[1] - scb[4]) * (scb[4] - scb[0]) * (scb[4] - scb[2]));
        pi[0][1] = ((scb[1] - scb[0]) * (scb[1] - scb[2]) *
                    ((scb[1] - scb[3]) * (scb[1] - scb[3]) - (scb[0] - scb[4]) * ((scb[3] - scb[1]) + (scb[4] - scb[1])))) /
                   ((scb[0] - scb[3]) * (scb[1] - scb[3]) * (scb[0] - scb[4]) * (scb[1] - scb[4]));
        pi[0][2] = (scb[1] - scb[0]) *
                   ((scb[1] - scb[2]) * (scb[1] - scb[3]) + ((scb[0] - scb[2]) + (scb[1] - scb[3])) * (scb[0] - scb[4])) /
                   ((scb[2] - scb[0]) * (scb[0] - scb[3]) * (scb[0] - scb[4]));
        pi[1][0] = ((scb[0] - scb[2]) * (scb[2] - scb[1]) * (scb[2] - scb[3])) /
                   ((scb[2] - scb[4]) * (scb[4] - scb[0]) * (scb[4] - scb[1]));
        pi[1][1] = ((scb[0] - scb[2]) * (scb[1] - scb[2]) *
                    ((scb[1] - scb[3]) * (scb[2] - scb[3]) + (scb[0] - scb[4]) * ((scb[1] - scb[3]) + (scb[2] - scb[4])))) /
                   ((scb[0] - scb[3]) * (scb[1] - scb[3]) * (scb[0] - scb[4]) * (scb[1] - scb[4]));
        pi[1][2] = ((scb[1] - scb[2]) * (scb[2] - scb[3]) * (scb[2] - scb[4])) /
                   ((scb[0] - scb[2]) * (scb[0] - scb[3]) * (scb[0] - scb[4]));
    }
}

static int is_cbc(int bc) { return bc <= MFC5_BC_CHAR_SLIP_WALL && bc >= MFC5_BC_CHAR_SUP_OUTFLOW; }

void orc_finalize_setup(orc_state* S) {
    for (int d = 0; d < S->nd; d++) {
        if (S->P.weno_order > 1) weno_coefficients(S, d);
        /* Del_in = Del_out = maxval(d<xyz>) over the ghosted array, m_cbc.fpp:312-313 */
        double mx = S->ds[d][-S->b];
        for (int i = -S->b; i <= S->s[d] + S->b; i++) mx = dmax(mx, S->ds[d][i]);
        S->Del[d] = mx;
        for (int side = 0; side < 2; side++)
            if (is_cbc(S->P.bc[d][side])) cbc_coefficients(S, d, side);
    }
}

/* --------------------------------------------------------------------------------------------- */
static void copy_interior(const orc_state* S, const double* src_packed, double* dst_field, int to_field) {
    int nx = S->s[0] + 1, ny = S->s[1] + 1, nz = S->s[2] + 1;
    long ncell = (long)nx * ny * nz;
    for (int v = 0; v < S->N; v++)
        for (int k = 0; k < nz; k++)
            for (int j = 0; j < ny; j++)
                for (int i = 0; i < nx; i++) {
                    long a = v * ncell + ((long)k * ny + j) * nx + i;
                    long f = v * S->fsz + IDX(S, i, j, k);
                    if (to_field) dst_field[f] = src_packed[a];
                    else ((double*)src_packed)[a] = dst_field[f];
                }
}
void orc_set_cons(orc_state* S, const double* q) { copy_interior(S, q, S->q1, 1); }
void orc_get_cons(const orc_state* S, double* q) { copy_interior(S, q, S->q1, 0); }
void orc_get_prim(const orc_state* S, double* q) { copy_interior(S, q, S->prim, 0); }
void orc_get_rhs(const orc_state* S, double* q) { copy_interior(S, q, S->rhs, 0); }
/* the buffered boxes themselves, sf(-b:m+b, -b:n+b, -b:p+b) per variable (collapsed dimensions 0:0), x fastest -- the layout
 * of q_cons_vf / q_prim_vf at the module boundary (m_time_steppers.fpp:179-185); which = 0 cons, 1 prim, 2 rhs */
long orc_box_size(const orc_state* S) { return S->fsz; }
void orc_get_box(const orc_state* S, int which, double* q) {
    const double* src = which == 0 ? S->q1 : which == 1 ? S->prim : S->rhs;
    memcpy(q, src, (size_t)S->N * S->fsz * sizeof(double));
}
void orc_set_cons_box(orc_state* S, const double* q) { memcpy(S->q1, q, (size_t)S->N * S->fsz * sizeof(double)); }

/* mixture rule shared by cons->prim, CBC and ICFL:
 * s_compute_species_fraction (m_variables_conversion.fpp:1140-1188) followed by
 * s_convert_species_to_mixture_variables_kernel (:198-270); with mpp_lim the clamp/renormalise is
 * applied by BOTH routines, in place. `passes` = 2 for callers that go through both. */
static void mixture(const orc_state* S, double* arho, double* alf, int passes, double* rho, double* gam, double* pinf, double* qv) {
    int nf = S->nf;
    if (S->P.mpp_lim) {
        for (int pass = 0; pass < passes; pass++) {
            double sum = 0.0;
            for (int i = 0; i < nf; i++) {
                arho[i] = dmax(0.0, arho[i]);
                alf[i] = dmin(dmax(0.0, alf[i]), 1.0);
                sum = sum + alf[i];
            }
            for (int i = 0; i < nf; i++) alf[i] = alf[i] / dmax(sum, SGM_EPS);
        }
    }
    double r = 0, g = 0, pi = 0, q = 0;
    for (int i = 0; i < nf; i++) {
        r = r + arho[i];
        g = g + alf[i] * S->P.fluid_gamma[i];
        pi = pi + alf[i] * S->P.fluid_pi_inf[i];
        q = q + arho[i] * S->P.fluid_qv[i];
    }
    *rho = r; *gam = g; *pinf = pi; *qv = q;
}

/* s_compute_speed_of_sound, 5-eq non-bubble branch: m_variables_conversion.fpp:1233-1260
 * (alt_soundspeed: Wood's mixture sound speed from the two bulk moduli, :1233-1236) */
static double sound_speed(const orc_state* S, double pres, double rho, double gam, double H, const double* adv, double vel_sum, double qv) {
    double c;
    if (S->P.alt_soundspeed) {
        double blkmod1 = ((S->P.fluid_gamma[0] + 1.0) * pres + S->P.fluid_pi_inf[0]) / S->P.fluid_gamma[0];
        double blkmod2 = ((S->P.fluid_gamma[1] + 1.0) * pres + S->P.fluid_pi_inf[1]) / S->P.fluid_gamma[1];
        c = (1.0 / (rho * (adv[0] / blkmod1 + adv[1] / blkmod2)));
    } else {
        c = (H - 5.e-1 * vel_sum - qv / rho) / gam;
    }
    if (S->P.mixture_err && c < 0.0) return 100.0 * SGM_EPS;
    return sqrt(c);
}

/* s_convert_conservative_to_primitive_variables: m_variables_conversion.fpp:459-717, on idwint
 * (m_rhs.fpp:641).  The simulation enables enforce_density_floor (m_start_up.fpp:823). */
void orc_cons2prim(orc_state* S) {
    int nf = S->nf, nd = S->nd;
#pragma omp parallel for collapse(2)
    for (int k = 0; k <= S->s[2]; k++)
        for (int j = 0; j <= S->s[1]; j++)
            for (int i = 0; i <= S->s[0]; i++) {
                long c = IDX(S, i, j, k);
                double arho[MAXF], alf[MAXF], rho, gam, pinf, qv;
                for (int f = 0; f < nf; f++) {
                    arho[f] = S->q1[f * S->fsz + c];
                    alf[f] = S->q1[(S->iadv + f) * S->fsz + c];
                }
                mixture(S, arho, alf, 2, &rho, &gam, &pinf, &qv);
                for (int f = 0; f < nf; f++) S->prim[f * S->fsz + c] = S->q1[f * S->fsz + c];
                rho = dmax(rho, SGM_EPS);
                double dyn = 0.0;
                for (int a = 0; a < nd; a++) {
                    long v = (S->imom + a) * S->fsz + c;
                    S->prim[v] = S->q1[v] / rho;
                    dyn = dyn + 5.e-1 * S->q1[v] * S->prim[v];
                }
                /* s_compute_pressure: m_variables_conversion.fpp:99 */
                S->prim[S->iE * S->fsz + c] = (S->q1[S->iE * S->fsz + c] - dyn - pinf - qv) / gam;
                for (int f = 0; f < nf; f++) S->prim[(S->iadv + f) * S->fsz + c] = S->q1[(S->iadv + f) * S->fsz + c];
            }
}

/* ---------------------------------------------------------------------------------------------
 * Ghost-cell fill of the primitive variables for one (direction, side):
 * s_populate_bc_direction m_boundary_common.fpp:120-187 with the per-cell kernels of
 * m_boundary_primitives.fpp: extrapolation :27-111 (codes -3..-12 and -17 without bc_io),
 * symmetry :114-335, periodic :338-494, slip wall :552-701, no-slip wall :704-874.
 * Transverse extents: x faces interior; y faces span x ghosts; z faces span x and y ghosts (:143-149).
 * ------------------------------------------------------------------------------------------- */
void orc_bc_fill(orc_state* S, int d, int side) {
    int bc = S->P.bc[d][side];
    if (bc >= 0) return;
    int b = S->b, s = S->s[d];
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    for (int e = 0; e < d; e++) { lo[e] = -S->off[e]; hi[e] = S->s[e] + S->off[e]; }
    lo[d] = hi[d] = 0;
    const double* wv = side == 0 ? S->P.bc_vb[d] : S->P.bc_ve[d];
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                int ijk[3] = {i, j, k};
                for (int l = 1; l <= b; l++) {
                    int g = side == 0 ? -l : s + l;             /* ghost index */
                    int edge = side == 0 ? 0 : s;               /* boundary cell */
                    int mir = side == 0 ? l - 1 : s - (l - 1);  /* mirror cell */
                    int per = side == 0 ? s - (l - 1) : l - 1;  /* periodic image */
                    ijk[d] = g; long cg = IDX(S, ijk[0], ijk[1], ijk[2]);
                    ijk[d] = edge; long ce = IDX(S, ijk[0], ijk[1], ijk[2]);
                    ijk[d] = mir; long cm = IDX(S, ijk[0], ijk[1], ijk[2]);
                    ijk[d] = per; long cp = IDX(S, ijk[0], ijk[1], ijk[2]);
                    for (int v = 0; v < S->N; v++) {
                        double* f = S->prim + v * S->fsz;
                        int a = v - S->imom; /* velocity component if 0 <= a < nd */
                        switch (bc) {
                        case MFC5_BC_PERIODIC: f[cg] = f[cp]; break;
                        case MFC5_BC_REFLECTIVE: f[cg] = (a == d) ? -f[cm] : f[cm]; break;
                        case MFC5_BC_SLIP_WALL: f[cg] = (a == d) ? -f[cm] + 2.0 * wv[d] : f[ce]; break;
                        case MFC5_BC_NO_SLIP_WALL: f[cg] = (a >= 0 && a < S->nd) ? -f[cm] + 2.0 * wv[a] : f[ce]; break;
                        case MFC5_BC_DIRICHLET:
                            if (S->dirich[d][side]) { /* s_dirichlet, m_boundary_primitives.fpp:877-969 */
                                int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
                                int n1 = S->s[t1] + 1, n2 = S->s[t2] + 1;
                                int a1 = ijk[t1], a2 = ijk[t2];
                                if (a1 < 0 || a1 >= n1 || a2 < 0 || a2 >= n2) f[cg] = 0.0; /* never packed by pre_process */
                                else f[cg] = S->dirich[d][side][((long)v * n2 + a2) * n1 + a1];
                            } else f[cg] = f[ce];
                            break;
                        default: f[cg] = f[ce]; break; /* -3 .. -12 */
                        }
                    }
                }
            }
}

/* Dirichlet buffers: what pre_process packs from the initial primitive field
 * (src/common/m_boundary_io.fpp:271-330) and the simulation reads back (bc_io). */
void orc_set_dirichlet(orc_state* S, int d, int side, const double* buf) {
    int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    long n = (long)S->N * (S->s[t1] + 1) * (S->s[t2] + 1);
    free(S->dirich[d][side]);
    S->dirich[d][side] = (double*)malloc(n * sizeof(double));
    memcpy(S->dirich[d][side], buf, n * sizeof(double));
}

/* Halo slabs for processor boundaries (bc >= 0): the b interior layers next to `side` are what the
 * neighbour on that side needs; extents as m_mpi_common.fpp:574-588,620-804 (y spans x ghosts,
 * z spans x and y ghosts).  Buffer order: var fastest is irrelevant here (both ends are ours):
 * (v, k, j, i) with i fastest. */
static void halo_bounds(const orc_state* S, int d, int lo[3], int hi[3]) {
    for (int e = 0; e < 3; e++) { lo[e] = 0; hi[e] = S->s[e]; }
    for (int e = 0; e < d; e++) { lo[e] = -S->off[e]; hi[e] = S->s[e] + S->off[e]; }
}
int orc_halo_count(const orc_state* S, int d) {
    int lo[3], hi[3]; halo_bounds(S, d, lo, hi);
    long n = S->N * (long)S->b;
    for (int e = 0; e < 3; e++) if (e != d) n *= (hi[e] - lo[e] + 1);
    return (int)n;
}
static void halo_copy(orc_state* S, int d, int first, double* buf, int pack) {
    int lo[3], hi[3]; halo_bounds(S, d, lo, hi);
    lo[d] = first; hi[d] = first + S->b - 1;
    long a = 0;
    for (int v = 0; v < S->N; v++)
        for (int k = lo[2]; k <= hi[2]; k++)
            for (int j = lo[1]; j <= hi[1]; j++)
                for (int i = lo[0]; i <= hi[0]; i++, a++) {
                    double* f = S->prim + v * S->fsz + IDX(S, i, j, k);
                    if (pack) buf[a] = *f; else *f = buf[a];
                }
}
void orc_halo_pack(const orc_state* S, int d, int side, double* buf) {
    halo_copy((orc_state*)S, d, side == 0 ? 0 : S->s[d] - S->b + 1, buf, 1);
}
void orc_halo_unpack(orc_state* S, int d, int side, const double* buf) {
    halo_copy(S, d, side == 0 ? -S->b : S->s[d] + 1, (double*)buf, 0);
}


/* ---------------------------------------------------------------------------------------------
 * MTHINC (int_comp = 2), src/simulation/m_thinc.fpp:28-291: multi-dimensional THINC with the interface normal from the
 * central difference of alpha_1 and the interface position d from a Newton solve of the cell's volume integral
 * (3-point Gauss-Legendre in the transverse directions, analytic along the first).
 * ------------------------------------------------------------------------------------------- */
static const double GQ3_PTS[3] = {-5e-1 * 0.7745966692414834, 0.0, 5e-1 * 0.7745966692414834};
static const double GQ3_WTS[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
#define VERYSMALL 1.e-12        /* m_constants.fpp:16 */
#define MTHINC_ALIGN_TOL 1.e-8  /* m_constants.fpp:56 */
static double log_cosh_diff(double a, double h) {
    double t = tanh(a) * tanh(h);
    return 2.0 * atanh(copysign(dmin(fabs(t), 1.0 - 2.220446049250313e-16), t));
}
static double thinc_integral_1d(double a, double bb) {
    if (fabs(bb) < VERYSMALL) return 5e-1 * (1.0 + tanh(a));
    return 5e-1 + log_cosh_diff(a, 5e-1 * bb) / (2.0 * bb);
}
static double mthinc_volume_integral(double n1, double n2, double n3, double dd, double beta, int ndim) {
    double res = 0.0;
    if (ndim == 1) return thinc_integral_1d(beta * dd, beta * n1);
    if (ndim == 2) {
        for (int q1 = 0; q1 < 3; q1++) res = res + GQ3_WTS[q1] * thinc_integral_1d(beta * (n2 * GQ3_PTS[q1] + dd), beta * n1);
        return res;
    }
    for (int q1 = 0; q1 < 3; q1++)
        for (int q2 = 0; q2 < 3; q2++)
            res = res + GQ3_WTS[q1] * GQ3_WTS[q2] * thinc_integral_1d(beta * (n2 * GQ3_PTS[q1] + n3 * GQ3_PTS[q2] + dd), beta * n1);
    return res;
}
static double mthinc_volume_integral_dd(double n1, double n2, double n3, double dd, double beta, int ndim) {
    double res = 0.0, th;
    if (ndim == 1) {
        for (int q1 = 0; q1 < 3; q1++) { th = tanh(beta * (n1 * GQ3_PTS[q1] + dd)); res = res + GQ3_WTS[q1] * (1.0 - th * th); }
    } else if (ndim == 2) {
        for (int q1 = 0; q1 < 3; q1++)
            for (int q2 = 0; q2 < 3; q2++) {
                th = tanh(beta * (n1 * GQ3_PTS[q1] + n2 * GQ3_PTS[q2] + dd));
                res = res + GQ3_WTS[q1] * GQ3_WTS[q2] * (1.0 - th * th);
            }
    } else {
        for (int q1 = 0; q1 < 3; q1++)
            for (int q2 = 0; q2 < 3; q2++)
                for (int q3 = 0; q3 < 3; q3++) {
                    th = tanh(beta * (n1 * GQ3_PTS[q1] + n2 * GQ3_PTS[q2] + n3 * GQ3_PTS[q3] + dd));
                    res = res + GQ3_WTS[q1] * GQ3_WTS[q2] * GQ3_WTS[q3] * (1.0 - th * th);
                }
    }
    return 5e-1 * beta * res;
}
static double mthinc_solve_d(double n1, double n2, double n3, double beta, double alpha_cell, int ndim) {
    double dd = 0.0;
    for (int iter = 1; iter <= 30; iter++) {
        double V = mthinc_volume_integral(n1, n2, n3, dd, beta, ndim);
        double residual = V - alpha_cell;
        if (fabs(residual) < VERYSMALL) break;
        double dV = mthinc_volume_integral_dd(n1, n2, n3, dd, beta, ndim);
        if (fabs(dV) < VERYSMALL) break;
        dd = dd - residual / dV;
    }
    return dd;
}
static double mthinc_face_average(double n1, double n2, double n3, double dd, double beta, int face_dir, double face_pos, int ndim) {
    double n_face, n_t0, n_t1, res;
    if (ndim == 1) return 5e-1 * (1.0 + tanh(beta * (n1 * face_pos + dd)));
    if (ndim == 2) {
        if (face_dir == 1) { n_face = n1; n_t0 = n2; } else { n_face = n2; n_t0 = n1; }
        return thinc_integral_1d(beta * (n_face * face_pos + dd), beta * n_t0);
    }
    if (face_dir == 1) { n_face = n1; n_t0 = n2; n_t1 = n3; }
    else if (face_dir == 2) { n_face = n2; n_t0 = n1; n_t1 = n3; }
    else { n_face = n3; n_t0 = n1; n_t1 = n2; }
    res = 0.0;
    for (int q = 0; q < 3; q++) res = res + GQ3_WTS[q] * thinc_integral_1d(beta * (n_face * face_pos + n_t0 * GQ3_PTS[q] + dd), beta * n_t1);
    return res;
}
/* s_compute_mthinc_normals, m_thinc.fpp:207-291; called once per RHS on the ghost-filled primitive variables (m_rhs.fpp:674) */
static void mthinc_normals(orc_state* S) {
    const int b = S->b, nd = S->nd;
    if (!S->mth_d) {
        for (int a = 0; a < 3; a++) S->mth_n[a] = (double*)calloc((size_t)S->fsz, sizeof(double));
        S->mth_d = (double*)calloc((size_t)S->fsz, sizeof(double));
    }
    for (int a = 0; a < 3; a++) memset(S->mth_n[a], 0, (size_t)S->fsz * sizeof(double));
    memset(S->mth_d, 0, (size_t)S->fsz * sizeof(double));
    /* cell centres with ghosts: x_cc(0:m) = x_cb(-1:m-1) + dx/2 (m_start_up.fpp), ghosts by s_populate_grid_variables_buffers */
    double* xcc[3] = {0, 0, 0};
    for (int d = 0; d < nd; d++) {
        const int s = S->s[d];
        double* st = (double*)calloc((size_t)(s + 1 + 2 * b), sizeof(double));
        double* cc = st + b;
        for (int i = 0; i <= s; i++) cc[i] = S->scb[d][i - 1] + S->ds[d][i] / 2.0;
        for (int i = 1; i <= b; i++) cc[-i] = cc[1 - i] - (S->ds[d][1 - i] + S->ds[d][-i]) / 2.0;
        for (int i = 1; i <= b; i++) cc[s + i] = cc[s + i - 1] + (S->ds[d][s + i - 1] + S->ds[d][s + i]) / 2.0;
        xcc[d] = cc;
    }
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < nd; d++) { lo[d] = -b + 1; hi[d] = S->s[d] + b - 1; }
    const double* al = S->prim + (long)S->iadv * S->fsz;
    const double ic_eps = S->P.ic_eps;
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                const long c = IDX(S, i, j, k);
                const double ac = al[c];
                if (!(ac >= ic_eps && ac <= 1.0 - ic_eps)) continue;
                const int ijk[3] = {i, j, k};
                double nr[3] = {0.0, 0.0, 0.0};
                for (int d = 0; d < nd; d++) {
                    const long sd = S->stride[d];
                    const int q = ijk[d];
                    nr[d] = (al[c + sd] - al[c - sd]) * (S->scb[d][q] - S->scb[d][q - 1]) / (xcc[d][q + 1] - xcc[d][q - 1]);
                }
                double nmag = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
                if (!(nmag > VERYSMALL)) continue;
                for (int d = 0; d < 3; d++) nr[d] = nr[d] / nmag;
                const double nmax = dmax(dmax(fabs(nr[0]), fabs(nr[1])), fabs(nr[2]));
                for (int d = 0; d < 3; d++) if (fabs(nr[d]) < MTHINC_ALIGN_TOL * nmax) nr[d] = 0.0;
                nmag = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
                if (!(nmag > VERYSMALL)) continue;
                for (int d = 0; d < 3; d++) { nr[d] = nr[d] / nmag; S->mth_n[d][c] = nr[d]; }
                S->mth_d[c] = mthinc_solve_d(nr[0], nr[1], nr[2], S->P.ic_beta, ac, nd);
            }
    for (int d = 0; d < nd; d++) free(xcc[d] - b);
}

/* WENO7 / TENO7 reconstruction of one variable in one cell, s_weno m_weno.fpp:1248-1490 (same operation order).
 * f points at v(i); sd = stride along the sweep.  Tables: cL/cR(sidx, r), bt(sidx, 0..5) with sidx = 0 the rightmost stencil;
 * dL/dR(q), poly(q), beta(q) numbered from the leftmost one. */
static void weno7_cell(const orc_state* S, int d, int jc, const double* f, long sd, double* oL, double* oR) {
    const int b = S->b, ns = S->ns, pn = 3;
    const double eps = S->P.weno_eps;
    const double* dl = &S->dL[d][(jc + b) * ns];
    const double* dr = &S->dR[d][(jc + b) * ns];
    const double* cl = &S->cL[d][(jc + b) * ns * pn];
    const double* cr = &S->cR[d][(jc + b) * ns * pn];
    const double* bt = &S->bt[d][(jc + b) * ns * 6];
    const double vp0 = f[0], vm1 = f[-sd], vm2 = f[-2 * sd], vm3 = f[-3 * sd], vp1 = f[sd], vp2 = f[2 * sd], vp3 = f[3 * sd];
    double poly[5], beta[5], alpha[5] = {0, 0, 0, 0, 0}, omega[5], delta[5] = {1, 1, 1, 1, 1}, tau = 0.0;
    if (!S->P.teno) {
        /* dvd(2) .. dvd(-3) */
        const double e2 = vp3 - vp2, e1 = vp2 - vp1, e0 = vp1 - vp0, em1 = vp0 - vm1, em2 = vm1 - vm2, em3 = vm2 - vm3;
        beta[3] = bt[0] * e2 * e2 + bt[1] * e2 * e1 + bt[2] * e2 * e0 + bt[3] * e1 * e1 + bt[4] * e1 * e0 + bt[5] * e0 * e0 + eps;
        beta[2] = bt[6] * e1 * e1 + bt[7] * e1 * e0 + bt[8] * e1 * em1 + bt[9] * e0 * e0 + bt[10] * e0 * em1 + bt[11] * em1 * em1 + eps;
        beta[1] = bt[12] * e0 * e0 + bt[13] * e0 * em1 + bt[14] * e0 * em2 + bt[15] * em1 * em1 + bt[16] * em1 * em2 + bt[17] * em2 * em2 + eps;
        beta[0] = bt[18] * em1 * em1 + bt[19] * em1 * em2 + bt[20] * em1 * em3 + bt[21] * em2 * em2 + bt[22] * em2 * em3 + bt[23] * em3 * em3 + eps;
        for (int side = 0; side < 2; side++) {
            const double* cc = side == 0 ? cl : cr;
            const double* dd = side == 0 ? dl : dr;
            poly[3] = vp0 + cc[0] * e2 + cc[1] * e1 + cc[2] * e0;
            poly[2] = vp0 + cc[3] * e1 + cc[4] * e0 + cc[5] * em1;
            poly[1] = vp0 + cc[6] * e0 + cc[7] * em1 + cc[8] * em2;
            poly[0] = vp0 + cc[9] * em1 + cc[10] * em2 + cc[11] * em3;
            if (S->P.mapped_weno) {
                for (int q = 0; q < 4; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
                double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3];
                for (int q = 0; q < 4; q++) omega[q] = alpha[q] / sum;
                for (int q = 0; q < 4; q++)
                    alpha[q] = (dd[q] * (1.0 + dd[q] - 3.0 * omega[q]) + omega[q] * omega[q]) *
                               (omega[q] / (dd[q] * dd[q] + omega[q] * (1.0 - 2.0 * dd[q])));
            } else if (S->P.wenoz) {
                if (side == 0) tau = fabs(beta[3] - beta[0]);
                for (int q = 0; q < 4; q++) alpha[q] = dd[q] * (1.0 + pow(tau / beta[q], S->P.wenoz_q));
            } else {
                for (int q = 0; q < 4; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
            }
            double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3];
            for (int q = 0; q < 4; q++) omega[q] = alpha[q] / sum;
            const double val = omega[0] * poly[0] + omega[1] * poly[1] + omega[2] * poly[2] + omega[3] * poly[3];
            if (side == 0) *oL = val; else *oR = val;
        }
        return;
    }
    /* TENO7 (Fu et al. 2016; Fu & Tang 2019): five uniform-grid stencils, m_weno.fpp:1303-1313,1355-1372,1427-1456 */
    static const double dLt[5] = {18.0 / 35.0, 3.0 / 35.0, 9.0 / 35.0, 1.0 / 35.0, 4.0 / 35.0};
    static const double dRt[5] = {18.0 / 35.0, 9.0 / 35.0, 3.0 / 35.0, 4.0 / 35.0, 1.0 / 35.0};
    beta[0] = 13.0 / 12.0 * sq(vm1 - 2.0 * vp0 + vp1) + sq(vm1 - vp1) / 4.0 + eps;
    beta[1] = 13.0 / 12.0 * sq(vp0 - 2.0 * vp1 + vp2) + sq(3.0 * vp0 - 4.0 * vp1 + vp2) / 4.0 + eps;
    beta[2] = 13.0 / 12.0 * sq(vm2 - 2.0 * vm1 + vp0) + sq(vm2 - 4.0 * vm1 + 3.0 * vp0) / 4.0 + eps;
    beta[3] = (vp0 * (2107.0 * vp0 - 9402.0 * vp1 + 7042.0 * vp2 - 1854.0 * vp3) + vp1 * (11003.0 * vp1 - 17246.0 * vp2 + 4642.0 * vp3) +
               vp2 * (7043.0 * vp2 - 3882.0 * vp3) + vp3 * (547.0 * vp3)) / 240.0 + eps;
    beta[4] = (vm3 * (547.0 * vm3 - 3882.0 * vm2 + 4642.0 * vm1 - 1854.0 * vp0) + vm2 * (7043.0 * vm2 - 17246.0 * vm1 + 7042.0 * vp0) +
               vm1 * (11003.0 * vm1 - 9402.0 * vp0) + vp0 * (2107.0 * vp0)) / 240.0 + eps;
    for (int side = 0; side < 2; side++) {
        const double* dd = side == 0 ? dLt : dRt;
        if (side == 0) {
            poly[0] = (2.0 * vm1 + 5.0 * vp0 - 1.0 * vp1) / 6.0;
            poly[1] = (11.0 * vp0 - 7.0 * vp1 + 2.0 * vp2) / 6.0;
            poly[2] = (-1.0 * vm2 + 5.0 * vm1 + 2.0 * vp0) / 6.0;
            poly[3] = (25.0 * vp0 - 23.0 * vp1 + 13.0 * vp2 - 3.0 * vp3) / 12.0;
            poly[4] = (1.0 * vm3 - 5.0 * vm2 + 13.0 * vm1 + 3.0 * vp0) / 12.0;
            tau = fabs(beta[4] - beta[3]);
            for (int q = 0; q < 5; q++) { alpha[q] = 1.0 + tau / beta[q]; alpha[q] = (alpha[q] * alpha[q] * alpha[q]) * (alpha[q] * alpha[q] * alpha[q]); }
            double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3] + alpha[4];
            for (int q = 0; q < 5; q++) { omega[q] = alpha[q] / sum; delta[q] = omega[q] < S->P.teno_CT ? 0.0 : 1.0; }
        } else {
            poly[0] = (-1.0 * vm1 + 5.0 * vp0 + 2.0 * vp1) / 6.0;
            poly[1] = (2.0 * vp0 + 5.0 * vp1 - 1.0 * vp2) / 6.0;
            poly[2] = (2.0 * vm2 - 7.0 * vm1 + 11.0 * vp0) / 6.0;
            poly[3] = (3.0 * vp0 + 13.0 * vp1 - 5.0 * vp2 + 1.0 * vp3) / 12.0;
            poly[4] = (-3.0 * vm3 + 13.0 * vm2 - 23.0 * vm1 + 25.0 * vp0) / 12.0;
        }
        for (int q = 0; q < 5; q++) alpha[q] = delta[q] * dd[q];
        double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3] + alpha[4];
        for (int q = 0; q < 5; q++) omega[q] = alpha[q] / sum;
        double val = omega[0] * poly[0] + omega[1] * poly[1] + omega[2] * poly[2] + omega[3] * poly[3];
        val = val + omega[4] * poly[4];
        if (side == 0) *oL = val; else *oR = val;
    }
}

/* ---------------------------------------------------------------------------------------------
 * WENO reconstruction along direction d: s_weno, m_weno.fpp:1007-1085 (WENO3), :1098-1239 (WENO5),
 * :952-993 (order 1).  vL = left-boundary value of cell j, vR = right-boundary value.  Only cells
 * -1..s+1 along the sweep and the transverse interior are consumed by the Riemann solver
 * (m_rhs.fpp:1020-1027), so only those are computed.
 * ------------------------------------------------------------------------------------------- */
static void weno_dir(orc_state* S, int d) {
    int b = S->b, pn = S->polyn, ns = S->ns;
    long sd = S->stride[d];
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    lo[d] = -1; hi[d] = S->s[d] + 1;
    const double eps = S->P.weno_eps;
    const int mapped = S->P.mapped_weno, wz = S->P.wenoz;
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                int ijk[3] = {i, j, k};
                int jc = ijk[d];
                long c = IDX(S, i, j, k);
                for (int v = 0; v < S->N; v++) {
                    const double* f = S->prim + v * S->fsz + c;
                    double* oL = S->vL + v * S->fsz + c;
                    double* oR = S->vR + v * S->fsz + c;
                    if (pn == 0) { *oL = f[0]; *oR = f[0]; continue; }
                    if (S->P.muscl_order > 0) { /* s_muscl, src/simulation/m_muscl.fpp:94-216 */
                        double slope = 0.0;
                        if (S->P.muscl_order == 2) {
                            const double meps = S->P.muscl_eps;
                            double slopeL = f[sd] - f[0], slopeR = f[0] - f[-sd];
                            switch (S->P.muscl_lim) {
                            case 0: slope = 5e-1 * (slopeL + slopeR); break;
                            case 1:
                                if (slopeL * slopeR > meps) slope = dmin(fabs(slopeL), fabs(slopeR));
                                if (slopeL < 0.0) slope = -slope;
                                break;
                            case 2:
                                if (slopeL * slopeR > meps) {
                                    slope = dmin(2.0 * fabs(slopeL), 2.0 * fabs(slopeR));
                                    slope = dmin(slope, 5e-1 * (fabs(slopeL) + fabs(slopeR)));
                                }
                                if (slopeL < 0.0) slope = -slope;
                                break;
                            case 3:
                                if (slopeL * slopeR > meps) slope = ((slopeL + slopeR) * slopeL * slopeR) / (slopeL * slopeL + slopeR * slopeR);
                                break;
                            case 4:
                                if (slopeL * slopeR > meps) slope = 2.0 * slopeL * slopeR / (slopeL + slopeR);
                                break;
                            case 5:
                                if (slopeL * slopeR > meps)
                                    slope = -1.0 * dmin(-dmin(2.0 * fabs(slopeL), fabs(slopeR)), -dmin(fabs(slopeL), 2.0 * fabs(slopeR)));
                                break;
                            }
                        }
                        *oL = f[0] - (5.e-1 * slope);
                        *oR = f[0] + (5.e-1 * slope);
                        continue;
                    }
                    if (pn == 3) { weno7_cell(S, d, jc, f, sd, oL, oR); continue; }
                    double dvd[4], poly[3], beta[3], alpha[3], omega[3], tau = 0.0, teno_delta[3] = {1.0, 1.0, 1.0};
                    const double* dl = &S->dL[d][(jc + b) * ns];
                    const double* dr = &S->dR[d][(jc + b) * ns];
                    const double* cl = &S->cL[d][(jc + b) * ns * pn];
                    const double* cr = &S->cR[d][(jc + b) * ns * pn];
                    const double* bt = &S->bt[d][(jc + b) * ns * 3];
                    double vp0 = f[0];
                    if (pn == 1) {
                        double vm1 = f[-sd], vp1 = f[sd];
                        double d0 = vp1 - vp0, dm1 = vp0 - vm1; /* dvd(0), dvd(-1) */
                        beta[0] = bt[0 * 3 + 0] * d0 * d0 + eps;
                        beta[1] = bt[1 * 3 + 0] * dm1 * dm1 + eps;
                        for (int sideLR = 0; sideLR < 2; sideLR++) {
                            const double* dd = sideLR == 0 ? dl : dr;
                            const double* cc = sideLR == 0 ? cl : cr;
                            poly[0] = vp0 + cc[0 * pn + 0] * d0;
                            poly[1] = vp0 + cc[1 * pn + 0] * dm1;
                            if (mapped) {
                                for (int q = 0; q < 2; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
                                double sum = alpha[0] + alpha[1];
                                for (int q = 0; q < 2; q++) omega[q] = alpha[q] / sum;
                                for (int q = 0; q < 2; q++)
                                    alpha[q] = (dd[q] * (1.0 + dd[q] - 3.0 * omega[q]) + omega[q] * omega[q]) *
                                               (omega[q] / (dd[q] * dd[q] + omega[q] * (1.0 - 2.0 * dd[q])));
                            } else if (wz) {
                                if (sideLR == 0) tau = fabs(beta[1] - beta[0]);
                                for (int q = 0; q < 2; q++) alpha[q] = dd[q] * (1.0 + tau / beta[q]);
                            } else {
                                for (int q = 0; q < 2; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
                            }
                            double sum = alpha[0] + alpha[1];
                            omega[0] = alpha[0] / sum; omega[1] = alpha[1] / sum;
                            double val = omega[0] * poly[0] + omega[1] * poly[1];
                            if (sideLR == 0) *oL = val; else *oR = val;
                        }
                    } else {
                        double vm1 = f[-sd], vm2 = f[-2 * sd], vp1 = f[sd], vp2 = f[2 * sd];
                        double d1 = vp2 - vp1, d0 = vp1 - vp0, dm1 = vp0 - vm1, dm2 = vm1 - vm2;
                        dvd[0] = d1; dvd[1] = d0; dvd[2] = dm1; dvd[3] = dm2;
                        (void)dvd;
                        if (S->uniform[d]) {
                            beta[0] = 13.0 / 12.0 * sq(d1 - d0) + 0.25 * sq(d1 - 3.0 * d0) + eps;
                            beta[1] = 13.0 / 12.0 * sq(d0 - dm1) + 0.25 * sq(d0 + dm1) + eps;
                            beta[2] = 13.0 / 12.0 * sq(dm1 - dm2) + 0.25 * sq(3.0 * dm1 - dm2) + eps;
                        } else {
                            beta[0] = bt[0] * d1 * d1 + bt[1] * d1 * d0 + bt[2] * d0 * d0 + eps;
                            beta[1] = bt[3] * d0 * d0 + bt[4] * d0 * dm1 + bt[5] * dm1 * dm1 + eps;
                            beta[2] = bt[6] * dm1 * dm1 + bt[7] * dm1 * dm2 + bt[8] * dm2 * dm2 + eps;
                        }
                        for (int sideLR = 0; sideLR < 2; sideLR++) {
                            const double* dd = sideLR == 0 ? dl : dr;
                            const double* cc = sideLR == 0 ? cl : cr;
                            poly[0] = vp0 + cc[0 * pn + 0] * d1 + cc[0 * pn + 1] * d0;
                            poly[1] = vp0 + cc[1 * pn + 0] * d0 + cc[1 * pn + 1] * dm1;
                            poly[2] = vp0 + cc[2 * pn + 0] * dm1 + cc[2 * pn + 1] * dm2;
                            if (mapped) {
                                for (int q = 0; q < 3; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
                                double sum = alpha[0] + alpha[1] + alpha[2];
                                for (int q = 0; q < 3; q++) omega[q] = alpha[q] / sum;
                                for (int q = 0; q < 3; q++)
                                    alpha[q] = (dd[q] * (1.0 + dd[q] - 3.0 * omega[q]) + omega[q] * omega[q]) *
                                               (omega[q] / (dd[q] * dd[q] + omega[q] * (1.0 - 2.0 * dd[q])));
                            } else if (wz) {
                                if (sideLR == 0) tau = fabs(beta[2] - beta[0]);
                                for (int q = 0; q < 3; q++) alpha[q] = dd[q] * (1.0 + (tau / beta[q]));
                            } else if (S->P.teno) { /* Fu et al. (2016), m_weno.fpp:1167-1187; the cut-off flags are reused on the right side (:1225-1229) */
                                if (sideLR == 0) {
                                    tau = fabs(beta[2] - beta[0]);
                                    for (int q = 0; q < 3; q++) {
                                        alpha[q] = 1.0 + tau / beta[q];
                                        alpha[q] = (alpha[q] * alpha[q] * alpha[q]) * (alpha[q] * alpha[q] * alpha[q]);
                                    }
                                    double sum = alpha[0] + alpha[1] + alpha[2];
                                    for (int q = 0; q < 3; q++) omega[q] = alpha[q] / sum;
                                    for (int q = 0; q < 3; q++) teno_delta[q] = omega[q] < S->P.teno_CT ? 0.0 : 1.0;
                                }
                                for (int q = 0; q < 3; q++) alpha[q] = teno_delta[q] * dd[q];
                            } else {
                                for (int q = 0; q < 3; q++) alpha[q] = dd[q] / (beta[q] * beta[q]);
                            }
                            omega[0] = alpha[0] / (alpha[0] + alpha[1] + alpha[2]);
                            omega[1] = alpha[1] / (alpha[0] + alpha[1] + alpha[2]);
                            omega[2] = alpha[2] / (alpha[0] + alpha[1] + alpha[2]);
                            double val = omega[0] * poly[0] + omega[1] * poly[1] + omega[2] * poly[2];
                            if (sideLR == 0) *oL = val; else *oR = val;
                        }
                        if (S->P.mp_weno) mp_limit(f[-2 * sd], f[-sd], f[0], f[sd], f[2 * sd], oL, oR); /* m_weno.fpp:1241-1243 */
                    }
                }
                /* interface compression (THINC), s_thinc_compression m_thinc.fpp:319-401: overwrites the reconstructed
                 * alpha / alpha-rho of the two fluids in interface cells where alpha_1 is monotone across the stencil */
                if (S->P.int_comp == 2 && S->nd > 1 && pn > 0) { /* MTHINC, m_thinc.fpp:323-355 */
                    const int nf = S->nf;
                    const double ic_eps = S->P.ic_eps;
                    const long ia = (long)S->iadv * S->fsz + c, ib = (long)(S->iadv + nf - 1) * S->fsz + c;
                    const long r0 = c, r1 = (long)(nf - 1) * S->fsz + c;
                    const double aC = S->prim[ia];
                    if (aC >= ic_eps && aC <= 1.0 - ic_eps) {
                        const double nh1 = S->mth_n[0][c], nh2 = S->mth_n[1][c], nh3 = S->mth_n[2][c], dloc = S->mth_d[c];
                        if (nh1 * nh1 + nh2 * nh2 + nh3 * nh3 > 5e-1) {
                            const double rho1 = S->prim[r0] / aC, rho2 = S->prim[r1] / (1.0 - aC);
                            double aT = mthinc_face_average(nh1, nh2, nh3, dloc, S->P.ic_beta, d + 1, -5e-1, S->nd);
                            if (aT < ic_eps) aT = ic_eps;
                            if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
                            S->vL[r0] = rho1 * aT; S->vL[r1] = rho2 * (1.0 - aT); S->vL[ia] = aT; S->vL[ib] = 1.0 - aT;
                            aT = mthinc_face_average(nh1, nh2, nh3, dloc, S->P.ic_beta, d + 1, 5e-1, S->nd);
                            if (aT < ic_eps) aT = ic_eps;
                            if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
                            S->vR[r0] = rho1 * aT; S->vR[r1] = rho2 * (1.0 - aT); S->vR[ia] = aT; S->vR[ib] = 1.0 - aT;
                        }
                    }
                } else if (S->P.int_comp >= 1 && pn > 0) { /* THINC (also int_comp = 2 in 1-D, m_thinc.fpp:324) */
                    const int nf = S->nf;
                    const double ic_eps = S->P.ic_eps, ic_beta = S->P.ic_beta;
                    const long ia = (long)S->iadv * S->fsz + c, ib = (long)(S->iadv + nf - 1) * S->fsz + c;
                    const long r0 = c, r1 = (long)(nf - 1) * S->fsz + c;
                    double aCL = S->prim[ia - sd], aC = S->prim[ia], aCR = S->prim[ia + sd];
                    if (aC >= ic_eps && aC <= 1.0 - ic_eps) {
                        double moncon = (aCR - aC) * (aC - aCL);
                        if (moncon > 1e-8) { /* moncon_cutoff, m_constants.fpp:55 */
                            double sgn = (aCR - aCL > 0.0) ? 1.0 : -1.0;
                            double beta_eff = ic_beta;
                            double qmin = dmin(aCR, aCL);
                            double qmax = dmax(aCR, aCL) - qmin;
                            double C = (aC - qmin + SGM_EPS) / (qmax + SGM_EPS);
                            double B = orc_exp(sgn * beta_eff * (2.0 * C - 1.0));
                            double A = (B / cosh(beta_eff) - 1.0) / tanh(beta_eff);
                            double rho_b = S->prim[r0] / aC;
                            double rho_e = S->prim[r1] / (1.0 - aC);
                            double aT = qmin + 5e-1 * qmax * (1.0 + sgn * A);
                            if (aT < ic_eps) aT = ic_eps;
                            if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
                            S->vL[r0] = rho_b * aT; S->vL[r1] = rho_e * (1.0 - aT); S->vL[ia] = aT; S->vL[ib] = 1.0 - aT;
                            aT = qmin + 5e-1 * qmax * (1.0 + sgn * (tanh(beta_eff) + A) / (1.0 + A * tanh(beta_eff)));
                            if (aT < ic_eps) aT = ic_eps;
                            if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
                            S->vR[r0] = rho_b * aT; S->vR[r1] = rho_e * (1.0 - aT); S->vR[ia] = aT; S->vR[ib] = 1.0 - aT;
                        }
                    }
                }
            }
}

/* ---------------------------------------------------------------------------------------------
 * Riemann solvers on faces -1..s of direction d (face j = right face of cell j):
 *   left state  = right-boundary value of cell j   (the solver's qL_prim_rs(j),   = S->vR[j])
 *   right state = left-boundary value of cell j+1  (the solver's qR_prim_rs(j+1), = S->vL[j+1])
 * (argument swap at m_rhs.fpp:1031-1033).  BC -4: m_riemann_state.fpp:149-206.
 * HLLC: m_riemann_solver_hllc.fpp:878-1332 (5-eq pure-fluid emission).
 * HLL : m_riemann_solver_hll.fpp:128-576.
 * Averages: include/inline_riemann.fpp:1-30.
 * ------------------------------------------------------------------------------------------- */
static void interface_reynolds(const orc_state* S, const double* alpha, double* Re);
static void riemann_dir(orc_state* S, int d) {
    int nf = S->nf, nd = S->nd;
    long sd = S->stride[d];
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    lo[d] = -1;
    const mfc5_params* P = &S->P;
    int didx[3];
    didx[0] = d;
    if (d == 0) { didx[1] = 1; didx[2] = 2; } else if (d == 1) { didx[1] = 0; didx[2] = 2; } else { didx[1] = 0; didx[2] = 1; }
    const int hllc = P->riemann_solver == 2;
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                int ijk[3] = {i, j, k};
                int jf = ijk[d];
                long c = IDX(S, i, j, k);
                double qL[MAXN], qR[MAXN];
                for (int v = 0; v < S->N; v++) {
                    qL[v] = S->vR[v * S->fsz + c];
                    qR[v] = S->vL[v * S->fsz + c + sd];
                }
                if (P->bc[d][0] == MFC5_BC_RIEMANN_EXTRAP && jf == -1)
                    for (int v = 0; v < S->N; v++) qL[v] = qR[v];
                if (P->bc[d][1] == MFC5_BC_RIEMANN_EXTRAP && jf == S->s[d])
                    for (int v = 0; v < S->N; v++) qR[v] = qL[v];

                double vel_L[3] = {0, 0, 0}, vel_R[3] = {0, 0, 0}, vel_L_rms = 0, vel_R_rms = 0;
                for (int a = 0; a < nd; a++) {
                    vel_L[a] = qL[S->imom + a]; vel_R[a] = qR[S->imom + a];
                    vel_L_rms = vel_L_rms + vel_L[a] * vel_L[a];
                    vel_R_rms = vel_R_rms + vel_R[a] * vel_R[a];
                }
                double pres_L = qL[S->iE], pres_R = qR[S->iE];
                double arho_L[MAXF], arho_R[MAXF], alf_L[MAXF], alf_R[MAXF];
                double adv_L[MAXF], adv_R[MAXF]; /* alpha_L/R handed to s_compute_speed_of_sound: hllc.fpp:889-892 (pre-limiter loads) */
                for (int f = 0; f < nf; f++) { adv_L[f] = qL[S->iadv + f]; adv_R[f] = qR[S->iadv + f]; }
                if (hllc) {
                    /* in-place limiter on the face buffers: hllc.fpp:938-958 */
                    if (P->mpp_lim) {
                        double sL = 0, sR = 0;
                        for (int f = 0; f < nf; f++) {
                            qL[f] = dmax(0.0, qL[f]);
                            qL[S->iadv + f] = dmin(dmax(0.0, qL[S->iadv + f]), 1.0);
                            qR[f] = dmax(0.0, qR[f]);
                            qR[S->iadv + f] = dmin(dmax(0.0, qR[S->iadv + f]), 1.0);
                            sL = sL + qL[S->iadv + f];
                            sR = sR + qR[S->iadv + f];
                        }
                        for (int f = 0; f < nf; f++) {
                            qL[S->iadv + f] = qL[S->iadv + f] / dmax(sL, SGM_EPS);
                            qR[S->iadv + f] = qR[S->iadv + f] / dmax(sR, SGM_EPS);
                        }
                    }
                    for (int f = 0; f < nf; f++) {
                        arho_L[f] = qL[f]; arho_R[f] = qR[f];
                        alf_L[f] = qL[S->iadv + f]; alf_R[f] = qR[S->iadv + f];
                    }
                } else {
                    /* HLL clamps local copies only: hll.fpp:187-200 */
                    for (int f = 0; f < nf; f++) {
                        arho_L[f] = qL[f]; arho_R[f] = qR[f];
                        alf_L[f] = qL[S->iadv + f]; alf_R[f] = qR[S->iadv + f];
                    }
                    if (P->mpp_lim) {
                        double sL = 0, sR = 0;
                        for (int f = 0; f < nf; f++) {
                            arho_L[f] = dmax(0.0, arho_L[f]);
                            alf_L[f] = dmin(dmax(0.0, alf_L[f]), 1.0);
                            sL = sL + alf_L[f];
                            arho_R[f] = dmax(0.0, arho_R[f]);
                            alf_R[f] = dmin(dmax(0.0, alf_R[f]), 1.0);
                            sR = sR + alf_R[f];
                        }
                        for (int f = 0; f < nf; f++) {
                            alf_L[f] = alf_L[f] / dmax(sL, SGM_EPS);
                            alf_R[f] = alf_R[f] / dmax(sR, SGM_EPS);
                        }
                    }
                }
                if (!hllc) for (int f = 0; f < nf; f++) { adv_L[f] = alf_L[f]; adv_R[f] = alf_R[f]; } /* hll.fpp:187-200: alpha_L itself is limited */
                /* s_accumulate_mixture_properties: m_riemann_state.fpp:1036-1058 */
                double rho_L = 0, gamma_L = 0, pi_inf_L = 0, qv_L = 0, rho_R = 0, gamma_R = 0, pi_inf_R = 0, qv_R = 0;
                for (int f = 0; f < nf; f++) {
                    rho_L = rho_L + arho_L[f];
                    gamma_L = gamma_L + alf_L[f] * P->fluid_gamma[f];
                    pi_inf_L = pi_inf_L + alf_L[f] * P->fluid_pi_inf[f];
                    qv_L = qv_L + arho_L[f] * P->fluid_qv[f];
                }
                for (int f = 0; f < nf; f++) {
                    rho_R = rho_R + arho_R[f];
                    gamma_R = gamma_R + alf_R[f] * P->fluid_gamma[f];
                    pi_inf_R = pi_inf_R + alf_R[f] * P->fluid_pi_inf[f];
                    qv_R = qv_R + arho_R[f] * P->fluid_qv[f];
                }
                double E_L = gamma_L * pres_L + pi_inf_L + 5.e-1 * rho_L * vel_L_rms + qv_L;
                double E_R = gamma_R * pres_R + pi_inf_R + 5.e-1 * rho_R * vel_R_rms + qv_R;
                double H_L = (E_L + pres_L) / rho_L;
                double H_R = (E_R + pres_R) / rho_R;
                /* average state */
                double rho_avg, vel_avg_rms, H_avg, gamma_avg, qv_avg;
                if (P->avg_state == 1) {
                    rho_avg = sqrt(rho_L * rho_R);
                    vel_avg_rms = 0.0;
                    for (int a = 0; a < nd; a++)
                        vel_avg_rms = vel_avg_rms + sq(sqrt(rho_L) * vel_L[a] + sqrt(rho_R) * vel_R[a]) / sq(sqrt(rho_L) + sqrt(rho_R));
                    H_avg = (sqrt(rho_L) * H_L + sqrt(rho_R) * H_R) / (sqrt(rho_L) + sqrt(rho_R));
                    gamma_avg = (sqrt(rho_L) * gamma_L + sqrt(rho_R) * gamma_R) / (sqrt(rho_L) + sqrt(rho_R));
                    vel_avg_rms = sq(sqrt(rho_L) * vel_L[0] + sqrt(rho_R) * vel_R[0]) / sq(sqrt(rho_L) + sqrt(rho_R));
                    qv_avg = (sqrt(rho_L) * qv_L + sqrt(rho_R) * qv_R) / (sqrt(rho_L) + sqrt(rho_R));
                } else {
                    rho_avg = 5.e-1 * (rho_L + rho_R);
                    vel_avg_rms = 0.0;
                    for (int a = 0; a < nd; a++) vel_avg_rms = vel_avg_rms + sq(5.e-1 * (vel_L[a] + vel_R[a]));
                    H_avg = 5.e-1 * (H_L + H_R);
                    gamma_avg = 5.e-1 * (gamma_L + gamma_R);
                    qv_avg = 5.e-1 * (qv_L + qv_R);
                }
                double c_L = sound_speed(S, pres_L, rho_L, gamma_L, H_L, adv_L, vel_L_rms, qv_L);
                double c_R = sound_speed(S, pres_R, rho_R, gamma_R, H_R, adv_R, vel_R_rms, qv_R);
                /* the average call passes pres_R and alpha_R as placeholders (hllc.fpp:1063-1066) */
                double c_avg = sound_speed(S, pres_R, rho_avg, gamma_avg, H_avg, adv_R, vel_avg_rms, qv_avg);
                /* low Mach corrections, inline_riemann.fpp:71-93 (call sites hllc.fpp:1084-1086,1137-1141; hll.fpp:400-404) */
                double zcoef = 0.0;
                if (P->low_Mach != 0) zcoef = dmin(1.0, dmax(sqrt(vel_L_rms) / c_L, sqrt(vel_R_rms) / c_R));
                if (hllc && P->low_Mach == 2) {
                    double vel_L_tmp = 5.e-1 * ((vel_L[d] + vel_R[d]) + zcoef * (vel_L[d] - vel_R[d]));
                    double vel_R_tmp = 5.e-1 * ((vel_L[d] + vel_R[d]) + zcoef * (vel_R[d] - vel_L[d]));
                    vel_L[d] = vel_L_tmp;
                    vel_R[d] = vel_R_tmp;
                }
                double uL = vel_L[d], uR = vel_R[d];
                double s_L, s_R, s_S;
                if (P->wave_speeds == 1) {
                    s_L = dmin(uL - c_L, uR - c_R);
                    s_R = dmax(uR + c_R, uL + c_L);
                    s_S = (pres_R - pres_L + rho_L * uL * (s_L - uL) - rho_R * uR * (s_R - uR)) / (rho_L * (s_L - uL) - rho_R * (s_R - uR));
                } else {
                    double pres_SL = 5.e-1 * (pres_L + pres_R + rho_avg * c_avg * (uL - uR));
                    double pres_SR = pres_SL;
                    double Ms_L = dmax(1.0, sqrt(1.0 + ((5.e-1 + gamma_L) / (1.0 + gamma_L)) * (pres_SL / pres_L - 1.0) * pres_L /
                                                           ((pres_L + pi_inf_L / (1.0 + gamma_L)))));
                    double Ms_R = dmax(1.0, sqrt(1.0 + ((5.e-1 + gamma_R) / (1.0 + gamma_R)) * (pres_SR / pres_R - 1.0) * pres_R /
                                                           ((pres_R + pi_inf_R / (1.0 + gamma_R)))));
                    s_L = uL - c_L * Ms_L;
                    s_R = uR + c_R * Ms_R;
                    s_S = 5.e-1 * ((uL + uR) + (pres_L - pres_R) / (rho_avg * c_avg));
                }
                double s_M = dmin(0.0, s_L), s_P = dmax(0.0, s_R);
                if (P->riemann_solver == 5) { /* Lax-Friedrichs: one symmetric speed, m_riemann_solver_lf.fpp:238-253 */
                    double vl = 0.0, vr = 0.0;
                    for (int a = 0; a < nd; a++) { vl = vl + vel_L[a] * vel_L[a]; vr = vr + vel_R[a] * vel_R[a]; }
                    vl = sqrt(vl); vr = sqrt(vr);
                    s_P = dmax(vl, vr) + dmax(c_L, c_R);
                    s_M = -s_P;
                    s_L = s_M; s_R = s_P;
                }
                double pcorr = 0.0;
                if (P->low_Mach == 1) {
                    if (hllc) pcorr = rho_L * rho_R * (s_L - uL) * (s_R - uR) * (uR - uL) / (rho_R * (s_R - uR) - rho_L * (s_L - uL)) * (zcoef - 1.0);
                    else pcorr = -(s_P - s_M) * (rho_L + rho_R) / 8.0 * (zcoef - 1.0);
                }
                double F[MAXN], Fsrc[MAXF], Vsrc[3] = {0.0, 0.0, 0.0};
                for (int f = 0; f < MAXF; f++) Fsrc[f] = 0.0;
                if (hllc) {
                    double xi_L = (s_L - uL) / dmin(s_L - s_S, -SGM_EPS);
                    double xi_R = (s_R - uR) / dmax(s_R - s_S, SGM_EPS);
                    double xi_L_m1 = (s_S - uL) / dmin(s_L - s_S, -SGM_EPS);
                    double xi_R_m1 = (s_S - uR) / dmax(s_R - s_S, SGM_EPS);
                    double xi_M = (5.e-1 + fsign(5.e-1, s_S));
                    double xi_P = (5.e-1 - fsign(5.e-1, s_S));
                    for (int f = 0; f < nf; f++)
                        F[f] = xi_M * qL[f] * (uL + s_M * xi_L_m1) + xi_P * qR[f] * (uR + s_P * xi_R_m1);
                    for (int ii = 0; ii < nd; ii++) {
                        int a = didx[ii];
                        if (a >= nd) continue;
                        double flg = (a == d) ? 1.0 : 0.0;
                        F[S->imom + a] = xi_M * (rho_L * (uL * vel_L[a] + s_M * (flg * s_L + (1.0 - flg) * vel_L[a]) * xi_L_m1) + flg * (pres_L)) +
                                         xi_P * (rho_R * (uR * vel_R[a] + s_P * (flg * s_R + (1.0 - flg) * vel_R[a]) * xi_R_m1) + flg * (pres_R)) +
                                         (s_M / s_L) * (s_P / s_R) * flg * pcorr;
                    }
                    F[S->iE] = xi_M * (uL * (E_L + pres_L) + s_M * (E_L * xi_L_m1 + xi_L * (s_S - uL) * (rho_L * s_S + pres_L / (s_L - uL)))) +
                               xi_P * (uR * (E_R + pres_R) + s_P * (E_R * xi_R_m1 + xi_R * (s_S - uR) * (rho_R * s_S + pres_R / (s_R - uR)))) +
                               (s_M / s_L) * (s_P / s_R) * pcorr * s_S;
                    for (int f = 0; f < nf; f++)
                        F[S->iadv + f] = xi_M * qL[S->iadv + f] * (uL + s_M * xi_L_m1) + xi_P * qR[S->iadv + f] * (uR + s_P * xi_R_m1);
                    /* vel_src for the normal direction -> flux_src(adv%beg): hllc.fpp:1319-1332 */
                    Fsrc[0] = xi_M * (uL + 1.0 * s_M * xi_L_m1) + xi_P * (uR + 1.0 * s_P * xi_R_m1);
                    for (int a = 0; a < nd; a++) {
                        double flg = (a == d) ? 1.0 : 0.0;
                        Vsrc[a] = xi_M * (vel_L[a] + flg * s_M * xi_L_m1) + xi_P * (vel_R[a] + flg * s_P * xi_R_m1);
                    }
                } else {
                    for (int f = 0; f < nf; f++)
                        F[f] = (s_M * arho_R[f] * uR - s_P * arho_L[f] * uL + s_M * s_P * (arho_L[f] - arho_R[f])) / (s_M - s_P);
                    for (int ii = 0; ii < nd; ii++) {
                        int a = didx[ii];
                        if (a >= nd) continue;
                        double flg = (a == d) ? 1.0 : 0.0;
                        F[S->imom + a] = (s_M * (rho_R * uR * vel_R[a] + flg * pres_R) - s_P * (rho_L * uL * vel_L[a] + flg * pres_L) +
                                          s_M * s_P * (rho_L * vel_L[a] - rho_R * vel_R[a])) / (s_M - s_P) +
                                         (s_M / s_L) * (s_P / s_R) * pcorr * (vel_R[a] - vel_L[a]);
                    }
                    F[S->iE] = (s_M * uR * (E_R + pres_R) - s_P * uL * (E_L + pres_L) + s_M * s_P * (E_L - E_R)) / (s_M - s_P) +
                               (s_M / s_L) * (s_P / s_R) * pcorr * (vel_R_rms - vel_L_rms) / 2.0;
                    /* Lax-Friedrichs always exports the interface alphas (m_riemann_solver_lf.fpp:306-313) although the RHS then
                     * reads flux_src(adv%beg) as a shared face velocity (adv_src_mode_vel_iface, m_global_parameters.fpp:903-906) */
                    if (!P->hll_u_interface || P->riemann_solver == 5) {
                        for (int f = 0; f < nf; f++) {
                            F[S->iadv + f] = (qL[S->iadv + f] - qR[S->iadv + f]) * s_M * s_P / (s_M - s_P);
                            Fsrc[f] = (s_M * qR[S->iadv + f] - s_P * qL[S->iadv + f]) / (s_M - s_P);
                        }
                    } else {
                        for (int f = 0; f < nf; f++) {
                            double aL = qL[S->iadv + f], aR = qR[S->iadv + f];
                            if (0.0 <= s_L) F[S->iadv + f] = aL * uL;
                            else if (s_R <= 0.0) F[S->iadv + f] = aR * uR;
                            else F[S->iadv + f] = (s_R * aL * uL - s_L * aR * uR + s_L * s_R * (aR - aL)) / (s_R - s_L);
                        }
                        if (0.0 <= s_L) Fsrc[0] = uL;
                        else if (s_R <= 0.0) Fsrc[0] = uR;
                        else Fsrc[0] = (s_R * uL - s_L * uR) / (s_R - s_L);
                    }
                }
                for (int v = 0; v < S->N; v++) S->flux[v * S->fsz + c] = F[v];
                for (int f = 0; f < nf; f++) S->fsrc[f * S->fsz + c] = Fsrc[f];
                if (P->viscous) {
                    /* Re_avg_rsx_vf: hllc.fpp:975-977,1073-1080 (alpha_L / alpha_R: the pre-limiter loads); hll.fpp:205-207,340-347 (its
                     * limited copies).  vel_src_rsx_vf: hllc.fpp:1317-1322; the HLL solver never writes it (zero as allocated) */
                    double Re_L[2], Re_R[2];
                    interface_reynolds(S, adv_L, Re_L);
                    interface_reynolds(S, adv_R, Re_R);
                    for (int q = 0; q < 2; q++) S->re_avg[q * S->fsz + c] = 2.0 / (1.0 / Re_L[q] + 1.0 / Re_R[q]);
                    for (int a = 0; a < nd; a++) S->vel_src[a * S->fsz + c] = Vsrc[a];
                }
            }
}

/* ---------------------------------------------------------------------------------------------
 * Viscous stresses, velocity gradients at the cell boundaries by finite differences (weno_Re_flux = F):
 * s_get_viscous m_viscous.fpp:515-830.  For the derivative along dd:
 *   normal faces   dqL_dd_n(dd)(c) = (v(c) - v(c - e_dd)) / (cc(c) - cc(c - 1)),   dqR_dd_n(dd)(c) = (v(c + e_dd) - v(c)) / (cc(c + 1) - cc(c))
 *   faces along n  dqL_dd_n(n)(c)  = 0.25 (L(c) + R(c) + L(c - e_n) + R(c - e_n)),  dqR_dd_n(n)(c)  = 0.25 (L(c + e_n) + R(c + e_n) + L(c) + R(c))
 * with L / R = dqL / dqR_dd_n(dd), over the reference's index ranges (the whole buffered box minus the cells whose stencil
 * would leave it).
 * ------------------------------------------------------------------------------------------- */
static void viscous_gradients_fd(orc_state* S) {
    const int nd = S->nd;
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    for (int e = 0; e < nd; e++) { lo[e] = -S->b; hi[e] = S->s[e] + S->b; }
    for (int dd = 0; dd < nd; dd++) {
        const long sd = S->stride[dd];
        const double* cc = S->scc[dd];
        for (int side = 0; side < 2; side++) {   /* 0: dqL (backward), 1: dqR (forward) */
            int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
            if (side == 0) l2[dd] = lo[dd] + 1; else h2[dd] = hi[dd] - 1;
            double* out = side == 0 ? S->dqL[dd][dd] : S->dqR[dd][dd];
#pragma omp parallel for collapse(2)
            for (int k = l2[2]; k <= h2[2]; k++)
                for (int j = l2[1]; j <= h2[1]; j++)
                    for (int i = l2[0]; i <= h2[0]; i++) {
                        const int ijk[3] = {i, j, k};
                        const int jc = ijk[dd];
                        const long c = IDX(S, i, j, k);
                        for (int a = 0; a < nd; a++) {
                            const double* v = S->prim + (S->imom + a) * S->fsz;
                            if (side == 0) out[a * S->fsz + c] = (v[c] - v[c - sd]) / (cc[jc] - cc[jc - 1]);
                            else out[a * S->fsz + c] = (v[c + sd] - v[c]) / (cc[jc + 1] - cc[jc]);
                        }
                    }
        }
    }
    for (int dd = 0; dd < nd; dd++)
        for (int n = 0; n < nd; n++) {
            if (n == dd) continue;
            const long sn = S->stride[n];
            const double *L = S->dqL[dd][dd], *R = S->dqR[dd][dd];
            for (int side = 0; side < 2; side++) {
                int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
                l2[dd] = lo[dd] + 1; h2[dd] = hi[dd] - 1;
                if (side == 0) l2[n] = lo[n] + 1; else h2[n] = hi[n] - 1;
                double* out = side == 0 ? S->dqL[dd][n] : S->dqR[dd][n];
#pragma omp parallel for collapse(2)
                for (int k = l2[2]; k <= h2[2]; k++)
                    for (int j = l2[1]; j <= h2[1]; j++)
                        for (int i = l2[0]; i <= h2[0]; i++) {
                            const long c = IDX(S, i, j, k);
                            for (int a = 0; a < nd; a++) {
                                const long o = a * S->fsz + c;
                                double t;
                                if (side == 0) t = (L[o] + R[o] + L[o - sn] + R[o - sn]);
                                else t = (L[o + sn] + R[o + sn] + L[o] + R[o]);
                                out[o] = 25.e-2 * t;
                            }
                        }
            }
        }
}

/* s_compute_interface_reynolds, m_riemann_state.fpp:1062-1090 */
static void interface_reynolds(const orc_state* S, const double* alpha, double* Re) {
    for (int i = 0; i < 2; i++) {
        Re[i] = -1.e6; /* dflt_real */
        if (S->re_size[i] > 0) Re[i] = 0.0;
        for (int q = 0; q < S->re_size[i]; q++) Re[i] = alpha[S->re_idx[i][q]] / S->res[i][q] + Re[i];
        Re[i] = 1.0 / dmax(Re[i], SGM_EPS);
    }
}

/* s_compute_cartesian_viscous_source_flux, m_riemann_state.fpp:805-975 (+ s_calculate_shear / bulk_stress_tensor :978-1031),
 * on faces -1..s of direction d after the Riemann solve of that direction.  The solver's "L" gradients are the RIGHT cell-boundary
 * fields of cell j and its "R" gradients the LEFT ones of cell j+1 (argument swap at m_rhs.fpp:1031-1033). */
static void viscous_source_flux(orc_state* S, int d) {
    const int nd = S->nd;
    const long sd = S->stride[d];
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    lo[d] = -1;
    const int shear = S->re_size[0] > 0, bulk = S->re_size[1] > 0;
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                const long c = IDX(S, i, j, k);
                double vga[3][3], tau[3][3];
                for (int a = 0; a < 3; a++) for (int e = 0; e < 3; e++) vga[a][e] = 0.0;
                for (int a = 0; a < nd; a++)
                    for (int dd = 0; dd < nd; dd++)
                        vga[a][dd] = 0.5 * (S->dqR[dd][d][a * S->fsz + c] + S->dqL[dd][d][a * S->fsz + c + sd]);
                double div = 0.0;
                for (int a = 0; a < nd; a++) div = div + vga[a][a];
                const double Re_shear = S->re_avg[c], Re_bulk = S->re_avg[S->fsz + c];
                double fs[4] = {0.0, 0.0, 0.0, 0.0}; /* flux_src of mom(1..nd), E; initialised to zero (s_initialize_riemann_solver :440-450) */
                if (shear) {
                    for (int a = 0; a < nd; a++)
                        for (int e = 0; e < nd; e++) {
                            tau[a][e] = (vga[e][a] + vga[a][e]) / Re_shear;
                            if (a == e) tau[a][e] = tau[a][e] - (2.0 / 3.0) * div / Re_shear;
                        }
                    for (int a = 0; a < nd; a++) {
                        fs[a] = fs[a] - tau[d][a];
                        fs[nd] = fs[nd] - S->vel_src[a * S->fsz + c] * tau[d][a];
                    }
                }
                if (bulk) {
                    for (int a = 0; a < nd; a++) for (int e = 0; e < nd; e++) tau[a][e] = 0.0;
                    for (int a = 0; a < nd; a++) tau[a][a] = div / Re_bulk;
                    for (int a = 0; a < nd; a++) {
                        fs[a] = fs[a] - tau[d][a];
                        fs[nd] = fs[nd] - S->vel_src[a * S->fsz + c] * tau[d][a];
                    }
                }
                for (int a = 0; a <= nd; a++) S->fvis[a * S->fsz + c] = fs[a];
            }
}

/* s_compute_additional_physics_rhs, m_rhs.fpp:1737-1960 (Cartesian, viscous): rhs(mom..E) += 1/ds (fvis(j-1) - fvis(j)) */
static void viscous_rhs(orc_state* S, int d) {
    const long sd = S->stride[d];
#pragma omp parallel for collapse(2)
    for (int k = 0; k <= S->s[2]; k++)
        for (int j = 0; j <= S->s[1]; j++)
            for (int i = 0; i <= S->s[0]; i++) {
                const int ijk[3] = {i, j, k};
                const long c = IDX(S, i, j, k);
                for (int a = 0; a <= S->nd; a++) {
                    const long v = (long)(S->imom + a) * S->fsz + c;
                    S->rhs[v] = S->rhs[v] + 1.0 / S->ds[d][ijk[d]] * (S->fvis[a * S->fsz + c - sd] - S->fvis[a * S->fsz + c]);
                }
            }
}

/* ---------------------------------------------------------------------------------------------
 * Characteristic boundary conditions on one (direction, side): s_cbc m_cbc.fpp:464-927 with
 * s_initialize_cbc :930-1182 (load + sign flips), s_finalize_cbc :1185-1345 (store),
 * s_convert_primitive_to_flux_variables m_variables_conversion.fpp:952-1135, wave amplitudes
 * m_compute_cbc.fpp:136-365.  Rewrites flux / flux_src on the boundary face and, for WENO3/5, the
 * first 1/2 interior faces.
 * ------------------------------------------------------------------------------------------- */
static void cbc_dir_side(orc_state* S, int d, int side) {
    const mfc5_params* P = &S->P;
    int nf = S->nf, nd = S->nd, N = S->N, b = S->b, s = S->s[d];
    int bc = P->bc[d][side];
    long sd = S->stride[d];
    const double sgn = side == 0 ? 1.0 : -1.0; /* sign(1, -cbc_loc) */
    const int hll1 = (P->riemann_solver == 1 && !P->hll_u_interface);
    const double* fd = S->fd[d][side];
    const double(*pi)[3] = S->pi[d][side];
    const double ds0 = S->cds[d][side][0];
    int lo[3] = {0, 0, 0}, hi[3] = {S->s[0], S->s[1], S->s[2]};
    lo[d] = hi[d] = 0;
    int didx[3];
    didx[0] = d;
    if (d == 0) { didx[1] = 1; didx[2] = 2; } else if (d == 1) { didx[1] = 0; didx[2] = 2; } else { didx[1] = 0; didx[2] = 1; }
    int nface = S->P.weno_order == 5 ? 2 : (S->P.weno_order == 3 ? 1 : 0); /* interpolated interior faces */
    for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
            for (int i = lo[0]; i <= hi[0]; i++) {
                int ijk[3] = {i, j, k};
                ijk[d] = 0;
                long base = IDX(S, ijk[0], ijk[1], ijk[2]); /* cell 0 along d */
                /* local layers jl = 0..b -> global cell (side==0 ? jl : s - jl) */
                double q[8][MAXN], Fc[8][MAXN], Fs[8][MAXF];
                for (int jl = 0; jl <= b; jl++) {
                    long c = base + sd * (side == 0 ? jl : s - jl);
                    for (int v = 0; v < N; v++) q[jl][v] = S->prim[v * S->fsz + c];
                    q[jl][S->imom + d] = q[jl][S->imom + d] * sgn;
                }
                /* local flux slots jl = -1..nface-1 -> global face (side==0 ? jl : s-1-jl) */
                double fl[4][MAXN], fsl[4][MAXF]; /* index jl+1 */
                if (nface == 0) {
                    long c = base + sd * (side == 0 ? 0 : s - 1);
                    for (int v = 0; v < N; v++) fl[1][v] = S->flux[v * S->fsz + c] * (v == S->imom + d ? 1.0 : sgn);
                    for (int f = 0; f < nf; f++) fsl[1][f] = S->fsrc[f * S->fsz + c] * (hll1 ? 1.0 : sgn);
                }
                /* cell-centre fluxes of the b+1 layers */
                for (int jl = 0; jl <= b && nface > 0; jl++) {
                    double arho[MAXF], alf[MAXF], vel[3] = {0, 0, 0}, vsum = 0, rho, gam, pinf, qv;
                    for (int f = 0; f < nf; f++) { arho[f] = q[jl][f]; alf[f] = q[jl][S->iadv + f]; }
                    for (int a = 0; a < nd; a++) { vel[a] = q[jl][S->imom + a]; }
                    for (int a = 0; a < nd; a++) vsum = vsum + vel[a] * vel[a];
                    double pres = q[jl][S->iE];
                    mixture(S, arho, alf, 1, &rho, &gam, &pinf, &qv);
                    double E = gam * pres + pinf + 5.e-1 * rho * vsum + qv;
                    for (int f = 0; f < nf; f++) Fc[jl][f] = arho[f] * vel[d];
                    for (int ii = 0; ii < nd; ii++) {
                        int a = didx[ii];
                        if (a >= nd) continue;
                        Fc[jl][S->imom + a] = rho * vel[d] * vel[a] + pres * (a == d ? 1.0 : 0.0);
                    }
                    Fc[jl][S->iE] = vel[d] * (E + pres);
                    if (hll1) {
                        for (int f = 0; f < nf; f++) { Fc[jl][S->iadv + f] = 0.0; Fs[jl][f] = alf[f]; }
                    } else {
                        for (int f = 0; f < nf; f++) { Fc[jl][S->iadv + f] = vel[d] * alf[f]; Fs[jl][f] = vel[d]; }
                    }
                }
                if (nface == 1) {
                    for (int v = 0; v < N; v++) fl[1][v] = Fc[0][v] + pi[0][0] * (Fc[1][v] - Fc[0][v]);
                    for (int f = 0; f < nf; f++) fsl[1][f] = Fs[0][f] + (Fs[1][f] - Fs[0][f]) * pi[0][0];
                } else if (nface == 2) {
                    for (int jl = 0; jl <= 1; jl++) {
                        for (int v = 0; v < N; v++)
                            fl[jl + 1][v] = Fc[jl][v] + pi[jl][0] * (Fc[3][v] - Fc[2][v]) + pi[jl][1] * (Fc[2][v] - Fc[1][v]) +
                                            pi[jl][2] * (Fc[1][v] - Fc[0][v]);
                        for (int f = 0; f < nf; f++)
                            fsl[jl + 1][f] = Fs[jl][f] + (Fs[3][f] - Fs[2][f]) * pi[jl][0] + (Fs[2][f] - Fs[1][f]) * pi[jl][1] +
                                             (Fs[1][f] - Fs[0][f]) * pi[jl][2];
                    }
                }
                /* FD of the primitive variables at layer 0 and the LODI system */
                double arho[MAXF], adv[MAXF], vel[3] = {0, 0, 0}, vel_K_sum = 0, rho, gam, pinf, qv;
                for (int f = 0; f < nf; f++) arho[f] = q[0][f];
                for (int a = 0; a < nd; a++) vel[a] = q[0][S->imom + a];
                for (int a = 0; a < nd; a++) vel_K_sum = vel_K_sum + vel[a] * vel[a];
                double pres = q[0][S->iE];
                for (int f = 0; f < nf; f++) adv[f] = q[0][S->iadv + f];
                mixture(S, arho, adv, 1, &rho, &gam, &pinf, &qv);
                double mf[MAXF];
                for (int f = 0; f < nf; f++) mf[f] = arho[f] / rho;
                double E = gam * pres + pinf + 5.e-1 * rho * vel_K_sum;
                double H = (E + pres) / rho;
                double c = sound_speed(S, pres, rho, gam, H, adv, vel_K_sum, qv);
                double darho_ds[MAXF], dvel_ds[3] = {0, 0, 0}, dpres_ds = 0, dadv_ds[MAXF];
                for (int f = 0; f < nf; f++) { darho_ds[f] = 0; dadv_ds[f] = 0; }
                for (int jl = 0; jl <= b; jl++) {
                    for (int f = 0; f < nf; f++) darho_ds[f] = q[jl][f] * fd[jl] + darho_ds[f];
                    for (int a = 0; a < nd; a++) dvel_ds[a] = q[jl][S->imom + a] * fd[jl] + dvel_ds[a];
                    dpres_ds = q[jl][S->iE] * fd[jl] + dpres_ds;
                    for (int f = 0; f < nf; f++) dadv_ds[f] = q[jl][S->iadv + f] * fd[jl] + dadv_ds[f];
                }
                double lambda[3] = {vel[d] - c, vel[d], vel[d] + c};
                double Ma = vel[d] / c;
                double L[MAXN];
                for (int v = 0; v < N; v++) L[v] = 0.0;
                const int last = N - 1;
                /* index map (0-based): L[0]; L[1..nf] density; L[nf+1..nf+nd-1] transverse velocity;
                 * L[iE..N-2] advection; L[N-1]. */
#define L1_BASE (lambda[0] * (dpres_ds - rho * c * dvel_ds[d]))
#define FILL_ALL(fac, l2)                                                                              \
    do {                                                                                               \
        for (int f = 0; f < nf; f++) L[1 + f] = (fac) * (l2) * (c * c * darho_ds[f] - mf[f] * dpres_ds); \
        for (int ii = 1; ii < nd; ii++) L[S->imom + ii] = (fac) * (l2) * dvel_ds[didx[ii]];             \
        for (int f = 0; f < nf; f++) L[S->iE + f] = (fac) * (l2) * dadv_ds[f];                          \
    } while (0)
                switch (bc) {
                case MFC5_BC_CHAR_SLIP_WALL:
                    L[0] = L1_BASE; L[last] = L[0]; break;
                case MFC5_BC_CHAR_NR_SUB_BUFFER: {
                    double lf = (5.e-1 - 5.e-1 * fsign(1.0, lambda[0]));
                    L[0] = lf * lambda[0] * (dpres_ds - rho * c * dvel_ds[d]);
                    lf = (5.e-1 - 5.e-1 * fsign(1.0, lambda[1]));
                    FILL_ALL(lf, lambda[1]);
                    lf = (5.e-1 - 5.e-1 * fsign(1.0, lambda[2]));
                    L[last] = lf * lambda[2] * (dpres_ds + rho * c * dvel_ds[d]);
                } break;
                case MFC5_BC_CHAR_NR_SUB_INFLOW:
                    L[0] = L1_BASE;
                    if (P->grcbc_in[d]) {
                        for (int f = 0; f < nf; f++)
                            L[1 + f] = pow(c, 3.0) * Ma * (arho[f] - P->alpha_rho_in[d][f]) / S->Del[d] - c * Ma * (pres - P->pres_in[d]) / S->Del[d];
                        if (nd > 1) {
                            L[S->imom + 1] = c * Ma * (vel[didx[1]] - P->vel_in[d][didx[1]]) / S->Del[d];
                            if (nd > 2) L[S->imom + 2] = c * Ma * (vel[didx[2]] - P->vel_in[d][didx[2]]) / S->Del[d];
                        }
                        for (int f = 0; f < nf; f++) L[S->iE + f] = c * Ma * (adv[f] - P->alpha_in[d][f]) / S->Del[d];
                        L[last] = rho * sq(c) * (1.0 + Ma) * (vel[d] + P->vel_in[d][d] * (side == 0 ? -1.0 : 1.0)) / S->Del[d] +
                                  c * (1.0 + Ma) * (pres - P->pres_in[d]) / S->Del[d];
                    }
                    break;
                case MFC5_BC_CHAR_NR_SUB_OUTFLOW:
                    L[0] = L1_BASE; FILL_ALL(1.0, lambda[1]); L[last] = 0.0;
                    if (P->grcbc_out[d]) {
                        L[last] = c * (1.0 - Ma) * (pres - P->pres_out[d]) / S->Del[d];
                        if (P->grcbc_vel_out[d])
                            L[last] = L[last] + rho * sq(c) * (1.0 - Ma) * (vel[d] + P->vel_out[d][d] * (side == 0 ? -1.0 : 1.0)) / S->Del[d];
                    }
                    break;
                case MFC5_BC_CHAR_FF_SUB_OUTFLOW:
                    L[0] = L1_BASE; FILL_ALL(1.0, lambda[1]); L[last] = L[0] + 2.0 * rho * c * lambda[1] * dvel_ds[d]; break;
                case MFC5_BC_CHAR_CP_SUB_OUTFLOW:
                    L[0] = L1_BASE; FILL_ALL(1.0, lambda[1]); L[last] = -L[0]; break;
                case MFC5_BC_CHAR_SUP_INFLOW: break;
                case MFC5_BC_CHAR_SUP_OUTFLOW:
                    L[0] = L1_BASE; FILL_ALL(1.0, lambda[1]); L[last] = lambda[2] * (dpres_ds + rho * c * dvel_ds[d]); break;
                }
                double dpres_dt = -5.e-1 * (L[last] + L[0]);
                double darho_dt[MAXF], dvel_dt[3] = {0, 0, 0}, dadv_dt[MAXF];
                for (int f = 0; f < nf; f++) darho_dt[f] = -(L[1 + f] - mf[f] * dpres_dt) / (c * c);
                for (int ii = 0; ii < nd; ii++) {
                    int a = didx[ii];
                    if (a >= nd) continue;
                    double flg = (a == d) ? 1.0 : 0.0;
                    dvel_dt[a] = flg * (L[0] - L[last]) / (2.0 * rho * c) + (flg - 1.0) * L[S->imom + ii];
                }
                double vel_dv_dt_sum = 0.0;
                for (int a = 0; a < nd; a++) vel_dv_dt_sum = vel_dv_dt_sum + vel[a] * dvel_dt[a];
                for (int f = 0; f < nf; f++) dadv_dt[f] = -L[S->iE + f];
                double drho_dt = 0, dgamma_dt = 0, dpi_inf_dt = 0, dqv_dt = 0;
                for (int f = 0; f < nf; f++) {
                    drho_dt = drho_dt + darho_dt[f];
                    dgamma_dt = dgamma_dt + dadv_dt[f] * P->fluid_gamma[f];
                    dpi_inf_dt = dpi_inf_dt + dadv_dt[f] * P->fluid_pi_inf[f];
                    dqv_dt = dqv_dt + darho_dt[f] * P->fluid_qv[f];
                }
                for (int f = 0; f < nf; f++) fl[0][f] = fl[1][f] + ds0 * darho_dt[f];
                for (int a = 0; a < nd; a++) fl[0][S->imom + a] = fl[1][S->imom + a] + ds0 * (vel[a] * drho_dt + rho * dvel_dt[a]);
                fl[0][S->iE] = fl[1][S->iE] + ds0 * (pres * dgamma_dt + gam * dpres_dt + dpi_inf_dt + dqv_dt + rho * vel_dv_dt_sum +
                                                    5.e-1 * drho_dt * vel_K_sum);
                if (hll1) {
                    for (int f = 0; f < nf; f++) fl[0][S->iadv + f] = 0.0;
                    for (int f = 0; f < nf; f++)
                        fsl[0][f] = 1.0 / dmax(fabs(vel[d]), SGM_EPS) * fsign(1.0, vel[d]) *
                                    (fl[1][S->iadv + f] + vel[d] * fsl[1][f] + ds0 * dadv_dt[f]);
                } else {
                    for (int f = 0; f < nf; f++) fl[0][S->iadv + f] = fl[1][S->iadv + f] + ds0 * dadv_dt[f];
                    for (int f = 0; f < nf; f++) fsl[0][f] = fsl[1][f];
                }
                /* store back with the sign conventions of s_finalize_cbc */
                for (int jl = -1; jl <= nface - 1; jl++) {
                    long cf = base + sd * (side == 0 ? jl : s - 1 - jl);
                    for (int v = 0; v < N; v++) S->flux[v * S->fsz + cf] = fl[jl + 1][v] * (v == S->imom + d ? 1.0 : sgn);
                    if (hll1) for (int f = 0; f < nf; f++) S->fsrc[f * S->fsz + cf] = fsl[jl + 1][f];
                    else S->fsrc[cf] = fsl[jl + 1][0] * sgn;
                }
            }
#undef L1_BASE
#undef FILL_ALL
}

/* flux difference + advection source: m_rhs.fpp:1107-1120 (x, assigns), :1183-1196 (y), :1345-1358 (z),
 * alpha source m_rhs.fpp:1440-1458 (HLL method 1), :1474-1488 (u-interface). */
static void flux_difference(orc_state* S, int d) {
    long sd = S->stride[d];
    int nf = S->nf;
    const int alpha_iface = (S->P.riemann_solver == 1 && !S->P.hll_u_interface);
#pragma omp parallel for collapse(2)
    for (int k = 0; k <= S->s[2]; k++)
        for (int j = 0; j <= S->s[1]; j++)
            for (int i = 0; i <= S->s[0]; i++) {
                int ijk[3] = {i, j, k};
                long c = IDX(S, i, j, k);
                double inv_ds = 1.0 / S->ds[d][ijk[d]];
                for (int v = 0; v < S->N; v++) {
                    double f1 = S->flux[v * S->fsz + c - sd], f2 = S->flux[v * S->fsz + c];
                    if (d == 0) S->rhs[v * S->fsz + c] = inv_ds * (f1 - f2);
                    else S->rhs[v * S->fsz + c] = S->rhs[v * S->fsz + c] + inv_ds * (f1 - f2);
                }
                double Kterm = 0.0;
                if (S->P.alt_soundspeed) { /* m_rhs.fpp:1070-1095: cell-centre pressure (prim) and volume fractions (cons) */
                    double pres = S->prim[S->iE * S->fsz + c];
                    double blkmod1 = ((S->P.fluid_gamma[0] + 1.0) * pres + S->P.fluid_pi_inf[0]) / S->P.fluid_gamma[0] + (4.0 / 3.0) * 0.0;
                    double blkmod2 = ((S->P.fluid_gamma[1] + 1.0) * pres + S->P.fluid_pi_inf[1]) / S->P.fluid_gamma[1] + (4.0 / 3.0) * 0.0;
                    double alpha1 = S->q1[S->iadv * S->fsz + c], alpha2 = S->q1[(S->iadv + nf - 1) * S->fsz + c];
                    Kterm = alpha1 * alpha2 * (blkmod2 - blkmod1) / (alpha1 * blkmod2 + alpha2 * blkmod1);
                }
                for (int f = 0; f < nf; f++) {
                    long v = (S->iadv + f) * S->fsz + c;
                    if (alpha_iface) {
                        double coeff = S->prim[(S->imom + d) * S->fsz + c];
                        double f1 = S->fsrc[f * S->fsz + c - sd], f2 = S->fsrc[f * S->fsz + c];
                        S->rhs[v] = S->rhs[v] + inv_ds * coeff * (f1 - f2);
                    } else {
                        double coeff = S->q1[v];
                        double f1 = S->fsrc[c], f2 = S->fsrc[c - sd];
                        S->rhs[v] = S->rhs[v] + inv_ds * coeff * (f1 - f2);
                    }
                }
                if (S->P.alt_soundspeed) { /* K div(u) correction, u-interface form: m_rhs.fpp:1489-1503,1568-1582,1714-1728 */
                    double f1 = S->fsrc[c], f2 = S->fsrc[c - sd];
                    long v0 = S->iadv * S->fsz + c, v1 = (S->iadv + nf - 1) * S->fsz + c;
                    S->rhs[v0] = S->rhs[v0] + Kterm * inv_ds * (f1 - f2);
                    S->rhs[v1] = S->rhs[v1] - Kterm * inv_ds * (f1 - f2);
                }
            }
}

void orc_sweeps(orc_state* S) {
    if (S->P.viscous) viscous_gradients_fd(S); /* s_get_viscous, m_rhs.fpp:658-664 */
    if (S->P.int_comp == 2 && S->nd > 1) mthinc_normals(S);
    for (int d = 0; d < S->nd; d++) {
        weno_dir(S, d);
        riemann_dir(S, d);
        if (S->P.viscous) viscous_source_flux(S, d); /* end of the Riemann solver, hllc.fpp:1548-1570 / hll.fpp:678-700 */
        for (int side = 0; side < 2; side++)
            if (is_cbc(S->P.bc[d][side])) cbc_dir_side(S, d, side);
        flux_difference(S, d);
        if (S->P.viscous) viscous_rhs(S, d); /* s_compute_additional_physics_rhs, m_rhs.fpp:729-734 */
    }
}

/* s_tvd_rk update: m_time_steppers.fpp:421-432 (coefficients), :486-507 (update) */
void orc_rk_update(orc_state* S, int stage, double dt) {
    static const double rk3[3][4] = {{1, 0, 1, 1}, {1, 3, 1, 4}, {2, 1, 2, 3}};
    static const double rk2[2][4] = {{1, 0, 1, 1}, {1, 1, 1, 2}};
    static const double rk1[1][4] = {{1, 0, 1, 1}};
    const double* rk = S->P.time_stepper == 3 ? rk3[stage - 1] : S->P.time_stepper == 2 ? rk2[stage - 1] : rk1[0];
    int nstage = S->P.time_stepper;
#pragma omp parallel for collapse(2)
    for (int v = 0; v < S->N; v++)
        for (int k = 0; k <= S->s[2]; k++)
            for (int j = 0; j <= S->s[1]; j++)
                for (int i = 0; i <= S->s[0]; i++) {
                    long a = v * S->fsz + IDX(S, i, j, k);
                    if (stage == 1 && nstage > 1) S->q2[a] = S->q1[a];
                    double q2 = nstage > 1 ? S->q2[a] : S->q1[a];
                    S->q1[a] = (rk[0] * S->q1[a] + rk[1] * q2 + rk[2] * dt * S->rhs[a]) / rk[3];
                }
    /* body forces, operator-split after the stage update: s_apply_bodyforces m_time_steppers.fpp:533,713-737 with
     * s_compute_body_forces_rhs m_body_forces.fpp:294-385 (acceleration at mytime, :206-218) */
    if (S->P.bf[0] || S->P.bf[1] || S->P.bf[2]) {
        double accel[3] = {0, 0, 0};
        for (int d = 0; d < 3; d++)
            if (S->P.bf[d]) accel[d] = S->P.bf_g[d] + S->P.bf_k[d] * sin(S->P.bf_w[d] * S->mytime - S->P.bf_p[d]);
        const double ldt = rk[2] * dt / rk[3];
#pragma omp parallel for collapse(2)
        for (int k = 0; k <= S->s[2]; k++)
            for (int j = 0; j <= S->s[1]; j++)
                for (int i = 0; i <= S->s[0]; i++) {
                    long c = IDX(S, i, j, k);
                    double rhoM = 0.0, rhs_m[3] = {0, 0, 0}, rhs_E = 0.0;
                    for (int f = 0; f < S->nf; f++) rhoM = rhoM + S->q1[f * S->fsz + c];
                    for (int d = 0; d < S->nd; d++)
                        if (S->P.bf[d]) {
                            rhs_m[d] = rhs_m[d] + rhoM * accel[d];
                            rhs_E = rhs_E + S->q1[(S->imom + d) * S->fsz + c] * accel[d];
                        }
                    for (int d = 0; d < S->nd; d++) S->q1[(S->imom + d) * S->fsz + c] = S->q1[(S->imom + d) * S->fsz + c] + ldt * rhs_m[d];
                    S->q1[S->iE * S->fsz + c] = S->q1[S->iE * S->fsz + c] + ldt * rhs_E;
                }
    }
}

void orc_set_time(orc_state* S, double t) { S->mytime = t; }

void orc_compute_rhs(orc_state* S) {
    orc_cons2prim(S);
    for (int d = 0; d < S->nd; d++) { orc_bc_fill(S, d, 0); orc_bc_fill(S, d, 1); }
    orc_sweeps(S);
}

void orc_tvd_rk_step(orc_state* S, double dt) {
    for (int st = 1; st <= S->P.time_stepper; st++) {
        orc_compute_rhs(S);
        orc_rk_update(S, st, dt);
    }
    S->mytime = S->mytime + dt; /* m_start_up.fpp:634 */
}

/* ICFL: m_data_output.fpp:159-224 with s_compute_enthalpy (m_sim_helpers.fpp:45-105) and
 * s_compute_stability_from_dt (:108-130); uses the primitive variables of the last RHS evaluation. */
double orc_icfl_max(orc_state* S, double dt) {
    int nf = S->nf, nd = S->nd;
    double mx = 0.0;
    for (int k = 0; k <= S->s[2]; k++)
        for (int j = 0; j <= S->s[1]; j++)
            for (int i = 0; i <= S->s[0]; i++) {
                long c = IDX(S, i, j, k);
                double arho[MAXF], alf[MAXF], rho, gam, pinf, qv, vel[3] = {0, 0, 0}, vs = 0;
                for (int f = 0; f < nf; f++) { arho[f] = S->prim[f * S->fsz + c]; alf[f] = S->prim[(S->iadv + f) * S->fsz + c]; }
                mixture(S, arho, alf, 2, &rho, &gam, &pinf, &qv);
                for (int a = 0; a < nd; a++) vel[a] = S->prim[(S->imom + a) * S->fsz + c];
                for (int a = 0; a < nd; a++) vs = vs + vel[a] * vel[a];
                double pres = S->prim[S->iE * S->fsz + c];
                double E = gam * pres + pinf + 5.e-1 * rho * vs + qv;
                double H = (E + pres) / rho;
                double cs = sound_speed(S, pres, rho, gam, H, alf, vs, qv);
                double icfl;
                if (nd == 3) icfl = dt / dmin(dmin(S->ds[0][i] / (fabs(vel[0]) + cs), S->ds[1][j] / (fabs(vel[1]) + cs)), S->ds[2][k] / (fabs(vel[2]) + cs));
                else if (nd == 2) icfl = dt / dmin(S->ds[0][i] / (fabs(vel[0]) + cs), S->ds[1][j] / (fabs(vel[1]) + cs));
                else icfl = (dt / S->ds[0][i]) * (fabs(vel[0]) + cs);
                if (icfl > mx || icfl != icfl) mx = icfl;
            }
    return mx;
}
