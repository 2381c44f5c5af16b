// mfc5_dev.cuh -- device-side numerics of the fused 5-equation sweep (sm_100a, FP64).
//
// Per-cell / per-face arithmetic written in the reference's operation order (file:line cited at
// each routine; paths relative to MFlowCode/MFC @ 0b234b19).  Everything is templated on the
// compile-time shape <NF fluids, ND dims, DIR sweep direction> so the per-thread state lives in
// registers; WENO order / variant / Riemann solver are template parameters where a specialised
// ("case-optimised") kernel exists and run-time (-1) in the generic kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/mfc5_params.h"

#define MFC5_SGM_EPS 1.e-16  // reference src/common/m_constants.fpp:16

namespace mfc5 {

struct DirTables {
    // WENO tables of one direction, index j in [-b+polyn, s+b-polyn] -> slot (j + b)
    // (reference src/simulation/m_weno.fpp:170-893)
    const double* cL;  // [slot][q][r]  poly_coef_cbL
    const double* cR;  // [slot][q][r]  poly_coef_cbR
    const double* dL;  // [slot][q]     d_cbL
    const double* dR;  // [slot][q]     d_cbR
    const double* bt;  // [slot][q][3]  beta_coef
    const double* ds;  // cell widths, index i in [-b, s+b] -> slot (i + b)
    const double* cc;  // cell centres, same indexing (viscous path only: finite-difference velocity gradients)
    // the same numbers as one row of MFC5_TAB doubles per slot (one base pointer, immediate offsets),
    // plus the ideal-weight terms of the WENO-M map; used by the division-light kernels:
    //   [0,6) cL  [6,12) cR  [12,15) dL  [15,18) dR  [18,27) beta
    //   [27,30) dL+dL^2  [30,33) dR+dR^2  [33,36) 1-2dL  [36,39) 1-2dR  [39,42) dL^2  [42,45) dR^2  [45] ds  [46] 1/ds
    const double* tab;
    const double* tabT;  // the same, transposed: [MFC5_TAB][nslot]
    // x kernel: per 30-cell warp segment, [segment][MFC5_TAB / 2][32 lanes][2]; lane l holds the row of cell
    // 30*segment - 1 + l (clamped to s+1), entries in adjacent pairs: lane-contiguous 16-byte loads at immediate offsets
    const double* tabX;
    int nslot;
    int uniform;       // uniform_grid(dir), m_weno.fpp:869-879
};
#define MFC5_TAB 48

// Everything a kernel needs, passed by value (__grid_constant__).
struct DevP {
    int s[3];          // m, n, p
    int nd, nf, N, b, pn, ns;
    int wo, wv, rs;    // weno_order; variant 0 JS / 1 M / 2 Z; riemann_solver 1 HLL / 2 HLLC
    int hll_u, avg, ws, mpp, mixerr, fast;
    int fast_riem;     // fast && wave_speeds = 1 && no low_Mach: what the case-optimised kernels (compile-time solver, EXT = false) cover
    int low_mach;      // 0 / 1 / 2, inline_riemann.fpp:71-93 (reference-order kernels only)
    int teno;          // TENO5 weights (m_weno.fpp:1167-1187), cut-off teno_ct (reference-order kernels only)
    double teno_ct;
    double wenoz_q;    // WENO-Z power at order 7 (m_weno.fpp:1381-1389)
    int mp_weno;       // monotonicity-preserving bounds on the WENO5 edge values (m_weno.fpp:1495-1601; reference-order kernels only)
    int muscl, muscl_lim;  // MUSCL reconstruction order (0 = WENO) and limiter, m_muscl.fpp (reference-order kernels only)
    double muscl_eps;
    int int_comp;      // 1: THINC interface compression after the reconstruction (m_thinc.fpp:319-401; reference-order kernels only)
    double ic_eps, ic_beta, ic_cosh, ic_tanh;  // cosh / tanh of ic_beta from the host
    int bf_on;         // body forces after the RK update of this stage (m_time_steppers.fpp:533): accelerations bf_a (0 where off)
    double bf_a[3], bf_ldt;  // bf_ldt = rk_coef(s,3) * dt / rk_coef(s,4)
    int alt_ss;        // alt_soundspeed: Wood sound speed + K div(u) term (reference-order kernels only)
    // viscous stresses (reference-order kernels only): Re_size / Re_idx / Res_gs of m_global_parameters.fpp:781-812
    int viscous, re_size[2], re_idx[2][MFC5_MAX_FLUIDS];
    double res[2][MFC5_MAX_FLUIDS];
    long long sy, sz, fs, base;  // strides (x stride = 1), field size, offset of cell (0,0,0)
    double eps;
    double gam[MFC5_MAX_FLUIDS], pinf[MFC5_MAX_FLUIDS], qv[MFC5_MAX_FLUIDS];
    DirTables T[3];
    int bc[3][2];
    double vb[3][3], ve[3][3];
    // characteristic BC replacement fluxes, [slot 0..nface][comp 0..N-1 flux, N.. src][plane]
    int cbc_on[3][2];
    int nface, ncbc_comp;
    double* cbcbuf[3][2];
    const double* dirich[3][2];
    // per launch
    const double* prim;
    double* primw;
    double* prim_next;  // != nullptr: the RK-fused last sweep also writes the primitive variables of the new state here
    const double* q1;   // conservative state read by this stage
    const double* q2;   // stage-1 copy (q_cons_ts(2))
    double* qout;       // where the fused RK update writes
    double* rhs;
    double rk0, rk1, rk2, rk3, irk3, dt;
    int do_rk;
    int chunk;          // marching chunk length (<= 0: chosen per launch)
    int xseg_lo, xseg_n;  // x kernel: this launch covers segments [xseg_lo, xseg_lo + xseg_n) (xseg_n = 0: all) ...
    int xseg_hole, xseg_hole_n;  // ... of the segment list with [xseg_hole, xseg_hole + xseg_hole_n) taken out (the boundary launch: both ends in one grid)
    int x_march;        // x sweep with the marching kernel (lanes along y) instead of the warp-shuffle kernel
    // Ghost modes [direction][side] (case-optimised kernels): how a sweep loader obtains the primitive variables of a ghost
    // cell.  MEM: from memory (filled by k_bc_faces or the halo exchange); CLAMP / MIRROR / WRAP: straight from the interior
    // cell the boundary condition copies (m_boundary_primitives.fpp:27-135 extrapolation -3, -5..-12; :114-335 reflective -2
    // with the normal velocity negated; :338-494 periodic -1), so that no ghost-fill kernel runs for those faces.
    int gmode[3][2];
};
#define MFC5_GM_MEM 0
#define MFC5_GM_CLAMP 1
#define MFC5_GM_MIRROR 2
#define MFC5_GM_WRAP 3

// index pl in [-b, s+b] along direction d -> the cell whose stored value the (ghost) cell pl holds under the ghost mode of its
// side; flip: the normal velocity changes sign (reflective wall).  Every mode is an affine map of pl, so a kernel computes the
// two (slope, offset, flip) triples once and the per-plane cost is a couple of selects, not a chain of branches.
struct GhostMap {
    int a_lo, b_lo, a_hi, b_hi, s;
    bool f_lo, f_hi;
    __device__ __forceinline__ GhostMap(const DevP& P, int d) {
        s = P.s[d];
        const int g0 = P.gmode[d][0], g1 = P.gmode[d][1];
        a_lo = g0 == MFC5_GM_CLAMP ? 0 : g0 == MFC5_GM_MIRROR ? -1 : 1;
        b_lo = g0 == MFC5_GM_MIRROR ? -1 : g0 == MFC5_GM_WRAP ? s + 1 : 0;
        a_hi = g1 == MFC5_GM_CLAMP ? 0 : g1 == MFC5_GM_MIRROR ? -1 : 1;
        b_hi = g1 == MFC5_GM_CLAMP ? s : g1 == MFC5_GM_MIRROR ? 2 * s + 1 : g1 == MFC5_GM_WRAP ? -s - 1 : 0;
        f_lo = g0 == MFC5_GM_MIRROR;
        f_hi = g1 == MFC5_GM_MIRROR;
    }
    __device__ __forceinline__ int src(int pl, bool& flip) const {
        const bool lo = pl < 0, hi = pl > s;
        flip = (lo & f_lo) | (hi & f_hi);
        const int a = lo ? a_lo : hi ? a_hi : 1, b = lo ? b_lo : hi ? b_hi : 0;
        return a * pl + b;
    }
};

__device__ __forceinline__ double sq(double x) { return x * x; }
__device__ __forceinline__ double fsign(double a, double b) { return copysign(fabs(a), b); }

// -------------------------------------------------------------------------------------------------
// Division-light formulations (product build).  Same algebra as the reference, different rounding:
// the nonlinear weights are only ever used normalised, omega_q = a_q / sum(a), so every a_q may be
// multiplied by a common positive factor.  Multiplying by the product of the denominators removes
// all per-stencil divisions; one reciprocal per side remains (two with the WENO-M map, whose argument
// must be a normalised weight).  The reference does 24 divisions per variable per cell for WENO5-M.
// -------------------------------------------------------------------------------------------------
// 1/x to double precision without the IEEE slow path: MUFU.RCP64H seed + two Newton steps.
// Valid for normal, finite x (sums of weights, densities, wave-speed differences).
__device__ __forceinline__ double rcp_nr(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    // e = 1 - x r;  r1 = r (1 + e) has residual exactly e^2, so the second Newton step is r1 (1 + e^2):
    // e^2 and r1 are independent -> 3 dependent FP64 operations after the seed instead of 4
    const double e = fma(-x, r, 1.0);
    const double r1 = fma(r, e, r);
    const double e2 = e * e;
    return fma(r1, e2, r1);
}
// t / d for a constant d whose correctly rounded reciprocal id is at hand: one multiplication plus one residual
// correction (Markstein) gives the correctly rounded quotient, i.e. the same bits as the reference's division.  The RK
// update divides by 3: a plain multiplication by 1/3 is one ulp off in a third of the cells, and one ulp of the total
// energy of a stiffened gas is 4e-7 in the pressure.
__device__ __forceinline__ double div_by_const(double t, double d, double id) {
    const double q = t * id;
    const double r = fma(-d, q, t);
    return fma(r, id, q);
}

// a / b with one reciprocal and a residual correction (Markstein): the correctly rounded quotient whenever r is within an
// ulp of 1/b, i.e. the reference's division bit for bit in all but rare cases, for 3 FP64 operations once r is at hand
__device__ __forceinline__ double div_r(double a, double b, double r) {
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// 1/a and 1/b from ONE reciprocal: r = 1/(a b), 1/a = r b, 1/b = r a  (8 FP64 operations instead of 10, one
// MUFU instead of two).  Callers guarantee that a b neither overflows nor underflows.
__device__ __forceinline__ void rcp_nr2(double a, double b, double& ia, double& ib) {
    const double r = rcp_nr(a * b);
    ia = r * b;
    ib = r * a;
}

// One cell's row of DirTables::tab.  RowAoS: MFC5_TAB consecutive doubles (marching kernels: staged in
// shared memory, block-uniform).  RowSoA: element k at p[k * stride] (x kernel: lanes own consecutive
// cells, so consecutive lanes read consecutive addresses of the transposed table).
struct RowAoS {
    static constexpr bool kAlwaysTable = false;  // uniform grids take the closed-form smoothness indicators
    static constexpr bool kHasMapTerms = true;   // read d + d^2, 1 - 2d, d^2 from the row
    const double* __restrict__ p;
    __device__ __forceinline__ double operator[](int k) const { return p[k]; }
};
// the same row for kernels that always evaluate the tabulated smoothness indicators (no run-time uniform-grid branch in the loop)
struct RowAoST {
    static constexpr bool kAlwaysTable = true;
    static constexpr bool kHasMapTerms = true;
    const double* __restrict__ p;
    __device__ __forceinline__ double operator[](int k) const { return p[k]; }
};
struct RowSoA {
    static constexpr bool kAlwaysTable = false;
    static constexpr bool kHasMapTerms = false;  // per-lane loads are dearer than 3 flops: recompute them
    const double* __restrict__ p;
    long long stride;
    __device__ __forceinline__ double operator[](int k) const { return __ldg(p + k * stride); }
};

struct RowSeg {
    static constexpr bool kAlwaysTable = false;
    static constexpr bool kHasMapTerms = true;
    const double* __restrict__ p;                // tabX + segment * MFC5_TAB * 32 + 2 * lane
    __device__ __forceinline__ double operator[](int k) const { return __ldg(p + (k >> 1) * 64 + (k & 1)); }
};

// WENO-M map of a normalised weight w with ideal weight d (m_weno.fpp:1150-1156), as numerator and
// denominator:  g(w) = w (d + d^2 - 3 d w + w^2) / (d^2 + w (1 - 2 d));  c1 = d + d^2, c2 = 1 - 2d, dsq = d^2
__device__ __forceinline__ void map_nd(double d, double c1, double c2, double dsq, double w, double& num, double& den) {
    num = w * fma(w, fma(-3.0, d, w), c1);
    den = fma(w, c2, dsq);
}

// NV variables of one cell at a time: every table value is loaded once and applied to all NV stencils
// (f[u][0..4] = v(j-2..j+2) of variable u), which also gives the FP64 pipe NV independent chains.
template <int WV, int NV, class ROW>
__device__ __forceinline__ void weno5_fast(const DevP& P, const DirTables& T, const ROW row, const double (&f)[NV][5],
                                           double (&eL)[NV], double (&eR)[NV]) {
    const double eps = P.eps;
    double d1[NV], d0[NV], dm1[NV], dm2[NV], g0[NV], g1[NV], g2[NV];
#pragma unroll
    for (int u = 0; u < NV; u++) { d1[u] = f[u][4] - f[u][3]; d0[u] = f[u][3] - f[u][2]; dm1[u] = f[u][2] - f[u][1]; dm2[u] = f[u][1] - f[u][0]; }
    {
        double b0[NV], b1[NV], b2[NV];
        if (!ROW::kAlwaysTable && T.uniform) {
#pragma unroll
            for (int u = 0; u < NV; u++) {
                b0[u] = 13.0 / 12.0 * sq(d1[u] - d0[u]) + 0.25 * sq(d1[u] - 3.0 * d0[u]) + eps;
                b1[u] = 13.0 / 12.0 * sq(d0[u] - dm1[u]) + 0.25 * sq(d0[u] + dm1[u]) + eps;
                b2[u] = 13.0 / 12.0 * sq(dm1[u] - dm2[u]) + 0.25 * sq(3.0 * dm1[u] - dm2[u]) + eps;
            }
        } else {  // B0 x^2 + B1 x y + B2 y^2 + eps  =  x (B0 x + B1 y) + (B2 y) y + eps : 5 operations
            const double r18 = row[18], r19 = row[19], r20 = row[20], r21 = row[21], r22 = row[22], r23 = row[23], r24 = row[24],
                         r25 = row[25], r26 = row[26];
#pragma unroll
            for (int u = 0; u < NV; u++) {
                b0[u] = fma(d1[u], fma(r18, d1[u], r19 * d0[u]), fma(r20 * d0[u], d0[u], eps));
                b1[u] = fma(d0[u], fma(r21, d0[u], r22 * dm1[u]), fma(r23 * dm1[u], dm1[u], eps));
                b2[u] = fma(dm1[u], fma(r24, dm1[u], r25 * dm2[u]), fma(r26 * dm2[u], dm2[u], eps));
            }
        }
        // common-denominator factors shared by both sides
#pragma unroll
        for (int u = 0; u < NV; u++) {
            if (WV == 2) {  // Z: a_q = d_q (1 + tau/b_q)  ~  d_q (b_q + tau) b_r b_s
                const double tau = fabs(b2[u] - b0[u]);
                g0[u] = (b0[u] + tau) * (b1[u] * b2[u]);
                g1[u] = (b1[u] + tau) * (b0[u] * b2[u]);
                g2[u] = (b2[u] + tau) * (b0[u] * b1[u]);
            } else {        // JS (and first stage of M): a_q = d_q / b_q^2  ~  d_q (b_r b_s)^2
                const double p12 = b1[u] * b2[u], p02 = b0[u] * b2[u], p01 = b0[u] * b1[u];
                g0[u] = p12 * p12; g1[u] = p02 * p02; g2[u] = p01 * p01;
            }
        }
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int cc = side * 6, dd = 12 + side * 3;
        const double dd0 = row[dd + 0], dd1 = row[dd + 1], dd2 = row[dd + 2];
        const double c0 = row[cc + 0], c1 = row[cc + 1], c2 = row[cc + 2], c3 = row[cc + 3], c4 = row[cc + 4], c5 = row[cc + 5];
        double m1[3], m2[3], m3[3];  // d + d^2, 1 - 2d, d^2
        if (WV == 1) {
            if (ROW::kHasMapTerms) {
#pragma unroll
                for (int q = 0; q < 3; q++) { m1[q] = row[27 + side * 3 + q]; m2[q] = row[33 + side * 3 + q]; m3[q] = row[39 + side * 3 + q]; }
            } else {
                m1[0] = fma(dd0, dd0, dd0); m2[0] = fma(-2.0, dd0, 1.0); m3[0] = dd0 * dd0;
                m1[1] = fma(dd1, dd1, dd1); m2[1] = fma(-2.0, dd1, 1.0); m3[1] = dd1 * dd1;
                m1[2] = fma(dd2, dd2, dd2); m2[2] = fma(-2.0, dd2, 1.0); m3[2] = dd2 * dd2;
            }
        }
#pragma unroll
        for (int u = 0; u < NV; u++) {
            // candidate polynomials as INCREMENTS over the cell value: the weighted increment is added to f2 at the
            // end, so a locally constant field is reconstructed exactly (the normalised sum f2 * sum(a) / sum(a) is
            // not, and its rounding error -- ~ ulp(p) on the pressure -- would enter the momentum flux difference)
            const double f2 = f[u][2];
            const double p0 = fma(c1, d0[u], c0 * d1[u]);
            const double p1 = fma(c3, dm1[u], c2 * d0[u]);
            const double p2 = fma(c5, dm2[u], c4 * dm1[u]);
            double a0 = dd0 * g0[u], a1 = dd1 * g1[u], a2 = dd2 * g2[u];
            if (WV == 1) {
                const double r = rcp_nr(a0 + a1 + a2);
                double n0, n1, n2, e0, e1, e2;
                map_nd(dd0, m1[0], m2[0], m3[0], a0 * r, n0, e0);
                map_nd(dd1, m1[1], m2[1], m3[1], a1 * r, n1, e1);
                map_nd(dd2, m1[2], m2[2], m3[2], a2 * r, n2, e2);
                a0 = n0 * (e1 * e2);
                a1 = n1 * (e0 * e2);
                a2 = n2 * (e0 * e1);
            }
            const double val = fma(fma(a2, p2, fma(a1, p1, a0 * p0)), rcp_nr(a0 + a1 + a2), f2);
            if (side == 0) eL[u] = val; else eR[u] = val;
        }
    }
}

template <int WV, int NV, class ROW>
__device__ __forceinline__ void weno3_fast(const DevP& P, const DirTables& T, const ROW row, const double (&f)[NV][5],
                                           double (&eL)[NV], double (&eR)[NV]) {
    const double eps = P.eps;
    double d0[NV], dm1[NV], g0[NV], g1[NV];
    const double r18 = row[18], r21 = row[21];
#pragma unroll
    for (int u = 0; u < NV; u++) {
        d0[u] = f[u][3] - f[u][2]; dm1[u] = f[u][2] - f[u][1];
        const double b0 = fma(r18 * d0[u], d0[u], eps);
        const double b1 = fma(r21 * dm1[u], dm1[u], eps);
        if (WV == 2) {
            const double tau = fabs(b1 - b0);
            g0[u] = (b0 + tau) * b1;
            g1[u] = (b1 + tau) * b0;
        } else {
            g0[u] = b1 * b1;
            g1[u] = b0 * b0;
        }
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int cc = side * 6, dd = 12 + side * 3;
        const double dd0 = row[dd + 0], dd1 = row[dd + 1];
        const double c0 = row[cc + 0], c1 = row[cc + 1];
        double m1[2], m2[2], m3[2];
        if (WV == 1) {
            if (ROW::kHasMapTerms) {
#pragma unroll
                for (int q = 0; q < 2; q++) { m1[q] = row[27 + side * 3 + q]; m2[q] = row[33 + side * 3 + q]; m3[q] = row[39 + side * 3 + q]; }
            } else {
                m1[0] = fma(dd0, dd0, dd0); m2[0] = fma(-2.0, dd0, 1.0); m3[0] = dd0 * dd0;
                m1[1] = fma(dd1, dd1, dd1); m2[1] = fma(-2.0, dd1, 1.0); m3[1] = dd1 * dd1;
            }
        }
#pragma unroll
        for (int u = 0; u < NV; u++) {
            const double f2 = f[u][2];
            const double p0 = c0 * d0[u];  // increments over the cell value, see weno5_fast
            const double p1 = c1 * dm1[u];
            double a0 = dd0 * g0[u], a1 = dd1 * g1[u];
            if (WV == 1) {
                const double r = rcp_nr(a0 + a1);
                double n0, n1, e0, e1;
                map_nd(dd0, m1[0], m2[0], m3[0], a0 * r, n0, e0);
                map_nd(dd1, m1[1], m2[1], m3[1], a1 * r, n1, e1);
                a0 = n0 * e1;
                a1 = n1 * e0;
            }
            const double val = fma(fma(a1, p1, a0 * p0), rcp_nr(a0 + a1), f2);
            if (side == 0) eL[u] = val; else eR[u] = val;
        }
    }
}

// Monotonicity-preserving bounds on the WENO5 edge values (mp_weno), s_preserve_monotonicity m_weno.fpp:1495-1601; same
// operation order as oracle/mfc5_oracle.c mp_limit (explicit comparisons, so that the sign of a zero is decided the same way).
__device__ __forceinline__ double mp_min(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double mp_max(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double mp_md(double a, double b) {
    const double s = copysign(1.0, 4.0 * a - b);
    const double mn = mp_min(mp_min(mp_min(fabs(4.0 * a - b), fabs(a)), fabs(4.0 * b - a)), fabs(b));
    return (s + copysign(1.0, 4.0 * b - a)) * fabs((s + copysign(1.0, a)) * (s + copysign(1.0, b))) * mn / 8.0;
}
__device__ __forceinline__ void mp_limit(double vm2, double vm1, double vp0, double vp1, double vp2, double& vL, double& vR) {
    const double alpha_mp = 2.0, beta_mp = 4.0 / 3.0;
    const double dm = vp0 + vm2 - vm1 * 2.0, d0 = vp1 + vm1 - vp0 * 2.0, dp = vp2 + vp0 - vp1 * 2.0;
    double d_MD = mp_md(dm, d0), d_LC = mp_md(d0, dp);
    double UL = vp0 - (vp1 - vp0) * alpha_mp;
    double MD = (vp0 + vm1 - d_MD) * 5.e-1;
    double LC = vp0 - (vp1 - vp0) * 5.e-1 + beta_mp * d_LC;
    double vmin = mp_max(mp_min(mp_min(vp0, vm1), MD), mp_min(mp_min(vp0, UL), LC));
    double vmax = mp_min(mp_max(mp_max(vp0, vm1), MD), mp_max(mp_max(vp0, UL), LC));
    vL = vL + (copysign(5.e-1, vmin - vL) + copysign(5.e-1, vmax - vL)) * mp_min(fabs(vmin - vL), fabs(vmax - vL));
    d_MD = mp_md(d0, dp); d_LC = mp_md(dm, d0);
    UL = vp0 + (vp0 - vm1) * alpha_mp;
    MD = (vp0 + vp1 - d_MD) * 5.e-1;
    LC = vp0 + (vp0 - vm1) * 5.e-1 + beta_mp * d_LC;
    vmin = mp_max(mp_min(mp_min(vp0, vp1), MD), mp_min(mp_min(vp0, UL), LC));
    vmax = mp_min(mp_max(mp_max(vp0, vp1), MD), mp_max(mp_max(vp0, UL), LC));
    vR = vR + (copysign(5.e-1, vmin - vR) + copysign(5.e-1, vmax - vR)) * mp_min(fabs(vmin - vR), fabs(vmax - vR));
}

// -------------------------------------------------------------------------------------------------
// WENO reconstruction of one variable in one cell: left- and right-boundary values.
// reference: s_weno, src/simulation/m_weno.fpp:1007-1085 (WENO3), :1098-1239 (WENO5), :952-993 (1).
// f[0..4] = v(j-2..j+2).  WO / WV < 0: read from P (generic kernels).
// -------------------------------------------------------------------------------------------------
// `row`: this cell's row of DirTables::tab (RowAoS / RowSoA), used when FAST
template <int WO, int WV, int FAST, class ROW>
__device__ __forceinline__ void weno_edges(const DevP& P, const DirTables& T, int slot, const ROW row, const double f0,
                                           const double f1, const double f2, const double f3, const double f4, double& eL, double& eR) {
    const int wo = (WO < 0) ? P.wo : WO;
    const int wv = (WV < 0) ? P.wv : WV;
    const double eps = P.eps;
    if (wo == 1) { eL = f2; eR = f2; return; }
    if (!FAST && WO < 0 && P.muscl > 0) {  // s_muscl, src/simulation/m_muscl.fpp:94-216 (generic kernels only)
        double slope = 0.0;
        if (P.muscl == 2) {
            const double meps = P.muscl_eps;
            const double slopeL = f3 - f2, slopeR = f2 - f1;
            switch (P.muscl_lim) {
            case 0: slope = 5e-1 * (slopeL + slopeR); break;
            case 1:
                if (slopeL * slopeR > meps) slope = fmin(fabs(slopeL), fabs(slopeR));
                if (slopeL < 0.0) slope = -slope;
                break;
            case 2:
                if (slopeL * slopeR > meps) {
                    slope = fmin(2.0 * fabs(slopeL), 2.0 * fabs(slopeR));
                    slope = fmin(slope, 5e-1 * (fabs(slopeL) + fabs(slopeR)));
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
                if (slopeL * slopeR > meps) slope = -1.0 * fmin(-fmin(2.0 * fabs(slopeL), fabs(slopeR)), -fmin(fabs(slopeL), 2.0 * fabs(slopeR)));
                break;
            }
        }
        eL = f2 - (5.e-1 * slope);
        eR = f2 + (5.e-1 * slope);
        return;
    }
    if (FAST) {
        const double f[1][5] = {{f0, f1, f2, f3, f4}};
        double l[1], r[1];
        if (wo == 3) {
            if (wv == 1) weno3_fast<1, 1>(P, T, row, f, l, r);
            else if (wv == 2) weno3_fast<2, 1>(P, T, row, f, l, r);
            else weno3_fast<0, 1>(P, T, row, f, l, r);
        } else {
            if (wv == 1) weno5_fast<1, 1>(P, T, row, f, l, r);
            else if (wv == 2) weno5_fast<2, 1>(P, T, row, f, l, r);
            else weno5_fast<0, 1>(P, T, row, f, l, r);
        }
        eL = l[0]; eR = r[0];
        return;
    }
    if (wo == 3) {
        const double* cl = T.cL + slot * 2;
        const double* cr = T.cR + slot * 2;
        const double* dl = T.dL + slot * 2;
        const double* dr = T.dR + slot * 2;
        const double* bt = T.bt + slot * 6;
        const double vp0 = f2;
        const double d0 = f3 - f2, dm1 = f2 - f1;
        const double b0 = __ldg(bt + 0) * d0 * d0 + eps;
        const double b1 = __ldg(bt + 3) * dm1 * dm1 + eps;
        double tau = 0.0;
        if (wv == 2) tau = fabs(b1 - b0);
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const double* dd = side == 0 ? dl : dr;
            const double* cc = side == 0 ? cl : cr;
            const double dd0 = dd[0], dd1 = dd[1];
            const double p0 = vp0 + cc[0] * d0;
            const double p1 = vp0 + cc[1] * dm1;
            double a0, a1;
            if (wv == 1) {
                a0 = dd0 / (b0 * b0);
                a1 = dd1 / (b1 * b1);
                const double sum = a0 + a1;
                const double w0 = a0 / sum, w1 = a1 / sum;
                a0 = (dd0 * (1.0 + dd0 - 3.0 * w0) + w0 * w0) * (w0 / (dd0 * dd0 + w0 * (1.0 - 2.0 * dd0)));
                a1 = (dd1 * (1.0 + dd1 - 3.0 * w1) + w1 * w1) * (w1 / (dd1 * dd1 + w1 * (1.0 - 2.0 * dd1)));
            } else if (wv == 2) {
                a0 = dd0 * (1.0 + tau / b0);
                a1 = dd1 * (1.0 + tau / b1);
            } else {
                a0 = dd0 / (b0 * b0);
                a1 = dd1 / (b1 * b1);
            }
            const double sum = a0 + a1;
            const double w0 = a0 / sum, w1 = a1 / sum;
            const double val = w0 * p0 + w1 * p1;
            if (side == 0) eL = val; else eR = val;
        }
        return;
    }
    // WENO5
    {
        const double* cl = T.cL + slot * 6;
        const double* cr = T.cR + slot * 6;
        const double* dl = T.dL + slot * 3;
        const double* dr = T.dR + slot * 3;
        const double* bt = T.bt + slot * 9;
        const double vp0 = f2;
        const double d1 = f4 - f3, d0 = f3 - f2, dm1 = f2 - f1, dm2 = f1 - f0;
        double b0, b1, b2;
        if (T.uniform) {
            b0 = 13.0 / 12.0 * sq(d1 - d0) + 0.25 * sq(d1 - 3.0 * d0) + eps;
            b1 = 13.0 / 12.0 * sq(d0 - dm1) + 0.25 * sq(d0 + dm1) + eps;
            b2 = 13.0 / 12.0 * sq(dm1 - dm2) + 0.25 * sq(3.0 * dm1 - dm2) + eps;
        } else {
            b0 = __ldg(bt + 0) * d1 * d1 + __ldg(bt + 1) * d1 * d0 + __ldg(bt + 2) * d0 * d0 + eps;
            b1 = __ldg(bt + 3) * d0 * d0 + __ldg(bt + 4) * d0 * dm1 + __ldg(bt + 5) * dm1 * dm1 + eps;
            b2 = __ldg(bt + 6) * dm1 * dm1 + __ldg(bt + 7) * dm1 * dm2 + __ldg(bt + 8) * dm2 * dm2 + eps;
        }
        double tau = 0.0;
        if (wv == 2) tau = fabs(b2 - b0);
        double td0 = 1.0, td1 = 1.0, td2 = 1.0;  // TENO cut-off flags, computed once and used on both sides
        const bool teno = (WV < 0) && P.teno;
        if (teno) {
            tau = fabs(b2 - b0);
            double g0 = 1.0 + tau / b0, g1 = 1.0 + tau / b1, g2 = 1.0 + tau / b2;
            g0 = (g0 * g0 * g0) * (g0 * g0 * g0); g1 = (g1 * g1 * g1) * (g1 * g1 * g1); g2 = (g2 * g2 * g2) * (g2 * g2 * g2);
            const double sum = g0 + g1 + g2;
            td0 = (g0 / sum < P.teno_ct) ? 0.0 : 1.0;
            td1 = (g1 / sum < P.teno_ct) ? 0.0 : 1.0;
            td2 = (g2 / sum < P.teno_ct) ? 0.0 : 1.0;
        }
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const double* dd = side == 0 ? dl : dr;
            const double* cc = side == 0 ? cl : cr;
            const double dd0 = dd[0], dd1 = dd[1], dd2 = dd[2];
            const double p0 = vp0 + cc[0] * d1 + cc[1] * d0;
            const double p1 = vp0 + cc[2] * d0 + cc[3] * dm1;
            const double p2 = vp0 + cc[4] * dm1 + cc[5] * dm2;
            double a0, a1, a2;
            if (wv == 1) {
                a0 = dd0 / (b0 * b0);
                a1 = dd1 / (b1 * b1);
                a2 = dd2 / (b2 * b2);
                const double sum = a0 + a1 + a2;
                const double w0 = a0 / sum, w1 = a1 / sum, w2 = a2 / sum;
                a0 = (dd0 * (1.0 + dd0 - 3.0 * w0) + w0 * w0) * (w0 / (dd0 * dd0 + w0 * (1.0 - 2.0 * dd0)));
                a1 = (dd1 * (1.0 + dd1 - 3.0 * w1) + w1 * w1) * (w1 / (dd1 * dd1 + w1 * (1.0 - 2.0 * dd1)));
                a2 = (dd2 * (1.0 + dd2 - 3.0 * w2) + w2 * w2) * (w2 / (dd2 * dd2 + w2 * (1.0 - 2.0 * dd2)));
            } else if (wv == 2) {
                a0 = dd0 * (1.0 + (tau / b0));
                a1 = dd1 * (1.0 + (tau / b1));
                a2 = dd2 * (1.0 + (tau / b2));
            } else if (teno) {
                a0 = td0 * dd0; a1 = td1 * dd1; a2 = td2 * dd2;
            } else {
                a0 = dd0 / (b0 * b0);
                a1 = dd1 / (b1 * b1);
                a2 = dd2 / (b2 * b2);
            }
            const double sum = a0 + a1 + a2;
            const double w0 = a0 / sum, w1 = a1 / sum, w2 = a2 / sum;
            const double val = w0 * p0 + w1 * p1 + w2 * p2;
            if (side == 0) eL = val; else eR = val;
        }
        if ((WV < 0) && P.mp_weno) mp_limit(f0, f1, f2, f3, f4, eL, eR);  // m_weno.fpp:1241-1243
    }
}

// WENO7 / TENO7 reconstruction of one variable in one cell (reference-order only): s_weno, m_weno.fpp:1248-1490.
// f[0..6] = v(j-3 .. j+3).  Tables: cL / cR(stencil, r) and beta(stencil, 0..5) with stencil 0 the rightmost one; the ideal
// weights d(q), the polynomials and the smoothness indicators are numbered from the leftmost one (m_weno.fpp:362-851).
// Same operation order as oracle/mfc5_oracle.c weno7_cell.
__device__ __forceinline__ void weno7_edges(const DevP& P, const DirTables& T, int slot, const double (&f)[7], double& eL, double& eR) {
    const double eps = P.eps;
    const double vm3 = f[0], vm2 = f[1], vm1 = f[2], vp0 = f[3], vp1 = f[4], vp2 = f[5], vp3 = f[6];
    double poly[5], beta[5], alpha[5] = {0, 0, 0, 0, 0}, omega[5], delta[5] = {1, 1, 1, 1, 1}, tau = 0.0;
    if (!P.teno) {
        const double* dl = T.dL + slot * 4;
        const double* dr = T.dR + slot * 4;
        const double* cl = T.cL + slot * 12;
        const double* cr = T.cR + slot * 12;
        const double* bt = T.bt + slot * 24;
        const double e2 = vp3 - vp2, e1 = vp2 - vp1, e0 = vp1 - vp0, em1 = vp0 - vm1, em2 = vm1 - vm2, em3 = vm2 - vm3;
        beta[3] = __ldg(bt + 0) * e2 * e2 + __ldg(bt + 1) * e2 * e1 + __ldg(bt + 2) * e2 * e0 + __ldg(bt + 3) * e1 * e1 + __ldg(bt + 4) * e1 * e0 + __ldg(bt + 5) * e0 * e0 + eps;
        beta[2] = __ldg(bt + 6) * e1 * e1 + __ldg(bt + 7) * e1 * e0 + __ldg(bt + 8) * e1 * em1 + __ldg(bt + 9) * e0 * e0 + __ldg(bt + 10) * e0 * em1 + __ldg(bt + 11) * em1 * em1 + eps;
        beta[1] = __ldg(bt + 12) * e0 * e0 + __ldg(bt + 13) * e0 * em1 + __ldg(bt + 14) * e0 * em2 + __ldg(bt + 15) * em1 * em1 + __ldg(bt + 16) * em1 * em2 + __ldg(bt + 17) * em2 * em2 + eps;
        beta[0] = __ldg(bt + 18) * em1 * em1 + __ldg(bt + 19) * em1 * em2 + __ldg(bt + 20) * em1 * em3 + __ldg(bt + 21) * em2 * em2 + __ldg(bt + 22) * em2 * em3 + __ldg(bt + 23) * em3 * em3 + eps;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const double* cc = side == 0 ? cl : cr;
            const double* dd = side == 0 ? dl : dr;
            poly[3] = vp0 + __ldg(cc + 0) * e2 + __ldg(cc + 1) * e1 + __ldg(cc + 2) * e0;
            poly[2] = vp0 + __ldg(cc + 3) * e1 + __ldg(cc + 4) * e0 + __ldg(cc + 5) * em1;
            poly[1] = vp0 + __ldg(cc + 6) * e0 + __ldg(cc + 7) * em1 + __ldg(cc + 8) * em2;
            poly[0] = vp0 + __ldg(cc + 9) * em1 + __ldg(cc + 10) * em2 + __ldg(cc + 11) * em3;
            double d4[4];
#pragma unroll
            for (int q = 0; q < 4; q++) d4[q] = __ldg(dd + q);
            if (P.wv == 1) {
#pragma unroll
                for (int q = 0; q < 4; q++) alpha[q] = d4[q] / (beta[q] * beta[q]);
                const double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3];
#pragma unroll
                for (int q = 0; q < 4; q++) omega[q] = alpha[q] / sum;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    alpha[q] = (d4[q] * (1.0 + d4[q] - 3.0 * omega[q]) + omega[q] * omega[q]) * (omega[q] / (d4[q] * d4[q] + omega[q] * (1.0 - 2.0 * d4[q])));
            } else if (P.wv == 2) {
                if (side == 0) tau = fabs(beta[3] - beta[0]);
#pragma unroll
                for (int q = 0; q < 4; q++) alpha[q] = d4[q] * (1.0 + pow(tau / beta[q], P.wenoz_q));
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) alpha[q] = d4[q] / (beta[q] * beta[q]);
            }
            const double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3];
#pragma unroll
            for (int q = 0; q < 4; q++) omega[q] = alpha[q] / sum;
            const double val = omega[0] * poly[0] + omega[1] * poly[1] + omega[2] * poly[2] + omega[3] * poly[3];
            if (side == 0) eL = val; else eR = val;
        }
        return;
    }
    // TENO7 (Fu et al. 2016; Fu & Tang 2019): five uniform-grid stencils, m_weno.fpp:1303-1313,1355-1372,1427-1456
    const double dLt[5] = {18.0 / 35.0, 3.0 / 35.0, 9.0 / 35.0, 1.0 / 35.0, 4.0 / 35.0};
    const double dRt[5] = {18.0 / 35.0, 9.0 / 35.0, 3.0 / 35.0, 4.0 / 35.0, 1.0 / 35.0};
    beta[0] = 13.0 / 12.0 * sq(vm1 - 2.0 * vp0 + vp1) + sq(vm1 - vp1) / 4.0 + eps;
    beta[1] = 13.0 / 12.0 * sq(vp0 - 2.0 * vp1 + vp2) + sq(3.0 * vp0 - 4.0 * vp1 + vp2) / 4.0 + eps;
    beta[2] = 13.0 / 12.0 * sq(vm2 - 2.0 * vm1 + vp0) + sq(vm2 - 4.0 * vm1 + 3.0 * vp0) / 4.0 + eps;
    beta[3] = (vp0 * (2107.0 * vp0 - 9402.0 * vp1 + 7042.0 * vp2 - 1854.0 * vp3) + vp1 * (11003.0 * vp1 - 17246.0 * vp2 + 4642.0 * vp3) +
               vp2 * (7043.0 * vp2 - 3882.0 * vp3) + vp3 * (547.0 * vp3)) / 240.0 + eps;
    beta[4] = (vm3 * (547.0 * vm3 - 3882.0 * vm2 + 4642.0 * vm1 - 1854.0 * vp0) + vm2 * (7043.0 * vm2 - 17246.0 * vm1 + 7042.0 * vp0) +
               vm1 * (11003.0 * vm1 - 9402.0 * vp0) + vp0 * (2107.0 * vp0)) / 240.0 + eps;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const double* dd = side == 0 ? dLt : dRt;
        if (side == 0) {
            poly[0] = (2.0 * vm1 + 5.0 * vp0 - 1.0 * vp1) / 6.0;
            poly[1] = (11.0 * vp0 - 7.0 * vp1 + 2.0 * vp2) / 6.0;
            poly[2] = (-1.0 * vm2 + 5.0 * vm1 + 2.0 * vp0) / 6.0;
            poly[3] = (25.0 * vp0 - 23.0 * vp1 + 13.0 * vp2 - 3.0 * vp3) / 12.0;
            poly[4] = (1.0 * vm3 - 5.0 * vm2 + 13.0 * vm1 + 3.0 * vp0) / 12.0;
            tau = fabs(beta[4] - beta[3]);
#pragma unroll
            for (int q = 0; q < 5; q++) { alpha[q] = 1.0 + tau / beta[q]; alpha[q] = (alpha[q] * alpha[q] * alpha[q]) * (alpha[q] * alpha[q] * alpha[q]); }
            const double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3] + alpha[4];
#pragma unroll
            for (int q = 0; q < 5; q++) { omega[q] = alpha[q] / sum; delta[q] = omega[q] < P.teno_ct ? 0.0 : 1.0; }
        } else {
            poly[0] = (-1.0 * vm1 + 5.0 * vp0 + 2.0 * vp1) / 6.0;
            poly[1] = (2.0 * vp0 + 5.0 * vp1 - 1.0 * vp2) / 6.0;
            poly[2] = (2.0 * vm2 - 7.0 * vm1 + 11.0 * vp0) / 6.0;
            poly[3] = (3.0 * vp0 + 13.0 * vp1 - 5.0 * vp2 + 1.0 * vp3) / 12.0;
            poly[4] = (-3.0 * vm3 + 13.0 * vm2 - 23.0 * vm1 + 25.0 * vp0) / 12.0;
        }
#pragma unroll
        for (int q = 0; q < 5; q++) alpha[q] = delta[q] * dd[q];
        const double sum = alpha[0] + alpha[1] + alpha[2] + alpha[3] + alpha[4];
#pragma unroll
        for (int q = 0; q < 5; q++) omega[q] = alpha[q] / sum;
        double val = omega[0] * poly[0] + omega[1] * poly[1] + omega[2] * poly[2] + omega[3] * poly[3];
        val = val + omega[4] * poly[4];
        if (side == 0) eL = val; else eR = val;
    }
}

// NV variables at once (the sweep kernels' rolled variable loop): FAST -> the shared-table versions above,
// otherwise the reference-order routine per variable.
template <int WO, int WV, int FAST, int NV, class ROW>
__device__ __forceinline__ void weno_edges_nv(const DevP& P, const DirTables& T, int slot, const ROW row, const double (&f)[NV][5],
                                              double (&eL)[NV], double (&eR)[NV]) {
    const int wo = (WO < 0) ? P.wo : WO;
    const int wv = (WV < 0) ? P.wv : WV;
    if (FAST && wo == 5) {
        if (wv == 1) weno5_fast<1, NV>(P, T, row, f, eL, eR);
        else if (wv == 2) weno5_fast<2, NV>(P, T, row, f, eL, eR);
        else weno5_fast<0, NV>(P, T, row, f, eL, eR);
        return;
    }
    if (FAST && wo == 3) {
        if (wv == 1) weno3_fast<1, NV>(P, T, row, f, eL, eR);
        else if (wv == 2) weno3_fast<2, NV>(P, T, row, f, eL, eR);
        else weno3_fast<0, NV>(P, T, row, f, eL, eR);
        return;
    }
#pragma unroll
    for (int u = 0; u < NV; u++) weno_edges<WO, WV, 0>(P, T, slot, row, f[u][0], f[u][1], f[u][2], f[u][3], f[u][4], eL[u], eR[u]);
}

// exp() of the THINC profile.  The strict build evaluates the same fdlibm-style sequence as the oracle (oracle/mfc5_oracle.c
// orc_exp: Cody-Waite reduction, degree-5 Remez polynomial, < 1 ulp) so that the two agree bit for bit; the product build uses
// CUDA's exp.
__device__ __forceinline__ double thinc_exp(double x) {
#ifdef MFC5_STRICT
    const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00,
                 P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (!(fabs(x) < 700.0)) return exp(x);
    const int k = (int)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
    const double hi = x - (double)k * ln2hi, lo = (double)k * ln2lo;
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    return ldexp(y, k);
#else
    return exp(x);
#endif
}

// THINC (s_thinc_compression, m_thinc.fpp:319-401) for one cell: aCL / aC / aCR = alpha_1 of cells j-1, j, j+1, rb / re =
// alpha-rho of the first / last fluid of cell j.  Returns false where the reconstruction is left alone; otherwise
// L[0..3] / R[0..3] = (alpha-rho_1, alpha-rho_NF, alpha_1, alpha_NF) on the left / right face of the cell.
__device__ __forceinline__ bool thinc_cell(const DevP& P, double aCL, double aC, double aCR, double rb, double re, double* L, double* R) {
    const double ic_eps = P.ic_eps;
    if (!(aC >= ic_eps && aC <= 1.0 - ic_eps)) return false;
    const double moncon = (aCR - aC) * (aC - aCL);
    if (!(moncon > 1e-8)) return false;  // moncon_cutoff, m_constants.fpp:55
    const double sgn = (aCR - aCL > 0.0) ? 1.0 : -1.0;
    const double beta_eff = P.ic_beta;
    const double qmin = fmin(aCR, aCL);
    const double qmax = fmax(aCR, aCL) - qmin;
    const double C = (aC - qmin + MFC5_SGM_EPS) / (qmax + MFC5_SGM_EPS);
    const double B = thinc_exp(sgn * beta_eff * (2.0 * C - 1.0));
    const double A = (B / P.ic_cosh - 1.0) / P.ic_tanh;
    const double rho_b = rb / aC;
    const double rho_e = re / (1.0 - aC);
    double aT = qmin + 5e-1 * qmax * (1.0 + sgn * A);
    if (aT < ic_eps) aT = ic_eps;
    if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
    L[0] = rho_b * aT; L[1] = rho_e * (1.0 - aT); L[2] = aT; L[3] = 1.0 - aT;
    aT = qmin + 5e-1 * qmax * (1.0 + sgn * (P.ic_tanh + A) / (1.0 + A * P.ic_tanh));
    if (aT < ic_eps) aT = ic_eps;
    if (aT > 1.0 - ic_eps) aT = 1.0 - ic_eps;
    R[0] = rho_b * aT; R[1] = rho_e * (1.0 - aT); R[2] = aT; R[3] = 1.0 - aT;
    return true;
}

// s_compute_speed_of_sound, 5-eq branch: src/common/m_variables_conversion.fpp:1233-1260
// (alt_soundspeed: Wood's mixture sound speed from the two bulk moduli; adv = the two volume fractions)
__device__ __forceinline__ double sound_speed(const DevP& P, double pres, double rho, double gam, double H, const double* adv,
                                              double vel_sum, double qv) {
    double c;
    if (P.alt_ss) {
        const double blkmod1 = ((P.gam[0] + 1.0) * pres + P.pinf[0]) / P.gam[0];
        const double blkmod2 = ((P.gam[1] + 1.0) * pres + P.pinf[1]) / P.gam[1];
        c = (1.0 / (rho * (adv[0] / blkmod1 + adv[1] / blkmod2)));
    } else {
        c = (H - 5.e-1 * vel_sum - qv / rho) / gam;
    }
    if (P.mixerr && c < 0.0) return 100.0 * MFC5_SGM_EPS;
    return sqrt(c);
}

// Kapila K coefficient of a cell (m_rhs.fpp:1070-1095): pressure from q_prim, volume fractions from q_cons
__device__ __forceinline__ double k_term(const DevP& P, double pres, double alpha1, double alpha2) {
    const double blkmod1 = ((P.gam[0] + 1.0) * pres + P.pinf[0]) / P.gam[0] + (4.0 / 3.0) * 0.0;
    const double blkmod2 = ((P.gam[1] + 1.0) * pres + P.pinf[1]) / P.gam[1] + (4.0 / 3.0) * 0.0;
    return alpha1 * alpha2 * (blkmod2 - blkmod1) / (alpha1 * blkmod2 + alpha2 * blkmod1);
}

// -------------------------------------------------------------------------------------------------
// Division-light Riemann solve (product build).
// Same formulas as riemann_face below; differences: reciprocals instead of divisions, and for HLLC only
// the star state on the upwind side of the contact is evaluated (the reference evaluates both and
// multiplies one by xi_M / xi_P = 0; identical unless the discarded side is non-finite).
// EXT (the generic kernels, run-time solver): also the pressure-velocity wave-speed estimates (wave_speeds = 2,
// hllc.fpp:1041-1066,1104-1119 / hll.fpp:352-396) and the low-Mach corrections (low_Mach = 1 / 2, inline_riemann.fpp:71-93);
// the case-optimised kernels are only dispatched for wave_speeds = 1 without low_Mach (DevP::fast_riem) and compile them out.
// -------------------------------------------------------------------------------------------------
// min / max by one comparison and a select (DSETP + 2 SEL).  fmin / fmax carry IEEE NaN handling that costs ~6 more
// integer-pipe instructions each, and every non-FP64 instruction takes an issue slot from the FP64 pipe;
// the operands here are never NaN unless the state already is.
#ifdef MFC5_IEEE_MINMAX
__device__ __forceinline__ double sel_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double sel_max(double a, double b) { return fmax(a, b); }
#else
// (inline PTX: nvcc turns the C++ conditional back into the IEEE max / min expansion)
__device__ __forceinline__ double sel_min(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double sel_max(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
#endif

template <int NF, int ND, int DIR, int RS, bool EXT = (RS < 0)>
__device__ __forceinline__ void riemann_face_fast(const DevP& P, double* qL, double* qR, double* F, double* Fs) {
    constexpr int IM = NF, IE = NF + ND, IA = NF + ND + 1;
    const int rs = (RS < 0) ? P.rs : RS;
    double vel_L_rms = 0.0, vel_R_rms = 0.0;
#pragma unroll
    for (int a = 0; a < ND; a++) {
        vel_L_rms = fma(qL[IM + a], qL[IM + a], vel_L_rms);
        vel_R_rms = fma(qR[IM + a], qR[IM + a], vel_R_rms);
    }
    const double pres_L = qL[IE], pres_R = qR[IE];
    double arho_L[NF], arho_R[NF], alf_L[NF], alf_R[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        arho_L[f] = qL[f]; arho_R[f] = qR[f];
        alf_L[f] = qL[IA + f]; alf_R[f] = qR[IA + f];
    }
    if (P.mpp) {
        double sL = 0.0, sR = 0.0;
#pragma unroll
        for (int f = 0; f < NF; f++) {
            arho_L[f] = sel_max(arho_L[f], 0.0);
            alf_L[f] = sel_min(sel_max(alf_L[f], 0.0), 1.0);
            sL = sL + alf_L[f];
            arho_R[f] = sel_max(arho_R[f], 0.0);
            alf_R[f] = sel_min(sel_max(alf_R[f], 0.0), 1.0);
            sR = sR + alf_R[f];
        }
        double iL, iR;
        rcp_nr2(sel_max(sL, MFC5_SGM_EPS), sel_max(sR, MFC5_SGM_EPS), iL, iR);
#pragma unroll
        for (int f = 0; f < NF; f++) { alf_L[f] = alf_L[f] * iL; alf_R[f] = alf_R[f] * iR; }
        if (rs == 2) {  // HLLC limits the face states in place (hllc.fpp:938-958); HLL only local copies
#pragma unroll
            for (int f = 0; f < NF; f++) { qL[f] = arho_L[f]; qR[f] = arho_R[f]; qL[IA + f] = alf_L[f]; qR[IA + f] = alf_R[f]; }
        }
    }
    double rho_L = 0, gamma_L = 0, pi_inf_L = 0, qv_L = 0, rho_R = 0, gamma_R = 0, pi_inf_R = 0, qv_R = 0;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        rho_L = rho_L + arho_L[f];
        gamma_L = fma(alf_L[f], P.gam[f], gamma_L);
        pi_inf_L = fma(alf_L[f], P.pinf[f], pi_inf_L);
        qv_L = fma(arho_L[f], P.qv[f], qv_L);
        rho_R = rho_R + arho_R[f];
        gamma_R = fma(alf_R[f], P.gam[f], gamma_R);
        pi_inf_R = fma(alf_R[f], P.pinf[f], pi_inf_R);
        qv_R = fma(arho_R[f], P.qv[f], qv_R);
    }
    const double E_L = gamma_L * pres_L + pi_inf_L + 5.e-1 * rho_L * vel_L_rms + qv_L;
    const double E_R = gamma_R * pres_R + pi_inf_R + 5.e-1 * rho_R * vel_R_rms + qv_R;
    double irho_L, irho_R, igam_L, igam_R;
    rcp_nr2(rho_L, rho_R, irho_L, irho_R);
    rcp_nr2(gamma_L, gamma_R, igam_L, igam_R);
    const double H_L = (E_L + pres_L) * irho_L;
    const double H_R = (E_R + pres_R) * irho_R;
    double c_L = (H_L - 5.e-1 * vel_L_rms - qv_L * irho_L) * igam_L;
    double c_R = (H_R - 5.e-1 * vel_R_rms - qv_R * irho_R) * igam_R;
    c_L = (P.mixerr && c_L < 0.0) ? 100.0 * MFC5_SGM_EPS : sqrt(c_L);
    c_R = (P.mixerr && c_R < 0.0) ? 100.0 * MFC5_SGM_EPS : sqrt(c_R);
    // low-Mach corrections (EXT): zcoef; low_Mach = 2 replaces the face-normal velocities of the HLLC solve's local copies
    double zcoef = 0.0;
    if (EXT && P.low_mach != 0) {
        double icL, icR;
        rcp_nr2(c_L, c_R, icL, icR);
        zcoef = sel_min(1.0, sel_max(sqrt(vel_L_rms) * icL, sqrt(vel_R_rms) * icR));
        if (rs == 2 && P.low_mach == 2) {
            const double a = qL[IM + DIR], b = qR[IM + DIR];
            qL[IM + DIR] = 5.e-1 * ((a + b) + zcoef * (a - b));
            qR[IM + DIR] = 5.e-1 * ((a + b) + zcoef * (b - a));
        }
    }
    const double uL = qL[IM + DIR], uR = qR[IM + DIR];
    double s_L, s_R, s_S_ws2 = 0.0;
    const bool ws2 = EXT && P.ws != 1;
    if (!ws2) {
        s_L = sel_min(uL - c_L, uR - c_R);
        s_R = sel_max(uR + c_R, uL + c_L);
    } else {   // pressure-velocity estimates; Roe (avg_state = 1) or arithmetic averages, inline_riemann.fpp:1-30
        double rho_avg, vel_avg_rms, H_avg, gamma_avg, qv_avg;
        if (P.avg == 1) {
            const double sl = sqrt(rho_L), sr = sqrt(rho_R);
            const double iw = rcp_nr(sl + sr);
            rho_avg = sqrt(rho_L * rho_R);
            H_avg = div_r(sl * H_L + sr * H_R, sl + sr, iw);
            gamma_avg = div_r(sl * gamma_L + sr * gamma_R, sl + sr, iw);
            const double w = sl + sr;
            vel_avg_rms = div_r(sq(sl * qL[IM] + sr * qR[IM]), w * w, rcp_nr(w * w));  // component 1 only, inline_riemann.fpp:28
            qv_avg = div_r(sl * qv_L + sr * qv_R, sl + sr, iw);
        } else {
            rho_avg = 5.e-1 * (rho_L + rho_R);
            vel_avg_rms = 0.0;
#pragma unroll
            for (int a = 0; a < ND; a++) vel_avg_rms = vel_avg_rms + sq(5.e-1 * (qL[IM + a] + qR[IM + a]));
            H_avg = 5.e-1 * (H_L + H_R);
            gamma_avg = 5.e-1 * (gamma_L + gamma_R);
            qv_avg = 5.e-1 * (qv_L + qv_R);
        }
        double ira, iga;
        rcp_nr2(rho_avg, gamma_avg, ira, iga);
        double c_avg = div_r(H_avg - 5.e-1 * vel_avg_rms - div_r(qv_avg, rho_avg, ira), gamma_avg, iga);
        c_avg = (P.mixerr && c_avg < 0.0) ? 100.0 * MFC5_SGM_EPS : sqrt(c_avg);
        const double pres_S = 5.e-1 * (pres_L + pres_R + rho_avg * c_avg * (uL - uR));
        const double gL1 = 1.0 + gamma_L, gR1 = 1.0 + gamma_R;
        double igL1, igR1, ipL, ipR;
        rcp_nr2(gL1, gR1, igL1, igR1);
        rcp_nr2(pres_L, pres_R, ipL, ipR);
        const double dL = pres_L + div_r(pi_inf_L, gL1, igL1), dR = pres_R + div_r(pi_inf_R, gR1, igR1);
        double idL, idR;
        rcp_nr2(dL, dR, idL, idR);
        const double Ms_L = sel_max(1.0, sqrt(1.0 + div_r(div_r(5.e-1 + gamma_L, gL1, igL1) * (div_r(pres_S, pres_L, ipL) - 1.0) * pres_L, dL, idL)));
        const double Ms_R = sel_max(1.0, sqrt(1.0 + div_r(div_r(5.e-1 + gamma_R, gR1, igR1) * (div_r(pres_S, pres_R, ipR) - 1.0) * pres_R, dR, idR)));
        s_L = uL - c_L * Ms_L;
        s_R = uR + c_R * Ms_R;
        const double rc = rho_avg * c_avg;
        s_S_ws2 = 5.e-1 * ((uL + uR) + div_r(pres_L - pres_R, rc, rcp_nr(rc)));
    }
    const double s_M = sel_min(s_L, 0.0), s_P = sel_max(s_R, 0.0);
    // low_Mach = 1: pressure correction of the momentum and energy fluxes; (s_M / s_L)(s_P / s_R) is 1 on subsonic faces, else 0
    double pcw = 0.0;
    if (EXT && P.low_mach == 1) {
        double pcorr;
        if (rs == 2) {
            const double den = rho_R * (s_R - uR) - rho_L * (s_L - uL);
            pcorr = div_r(rho_L * rho_R * (s_L - uL) * (s_R - uR) * (uR - uL), den, rcp_nr(den)) * (zcoef - 1.0);
        } else {
            pcorr = -(s_P - s_M) * (rho_L + rho_R) * 0.125 * (zcoef - 1.0);
        }
        pcw = (s_M / s_L) * (s_P / s_R) * pcorr;
    }
    if (rs == 2) {
        const double mL = rho_L * (s_L - uL), mR = rho_R * (s_R - uR);
        // s_S = A / B.  Its sign (which selects the upwind side) is known before the division, and
        // 1 / (s_K - s_S) = B / (s_K B - A): both reciprocals come out of one (rcp_nr2), in parallel
        // with the one of s_K - u_K, instead of two dependent ones.  (wave_speeds = 2: s_S is given.)
        const double A = pres_R - pres_L + mL * uL - mR * uR, B = mL - mR;
        const bool left = ws2 ? !signbit(s_S_ws2) : (signbit(A) == signbit(B));  // xi_M = 1, xi_P = 0  (Fortran sign(0.5, s_S))
        const double* qK = left ? qL : qR;
        const double rho_K = left ? rho_L : rho_R, u_K = left ? uL : uR, p_K = left ? pres_L : pres_R, E_K = left ? E_L : E_R;
        const double s_K = left ? s_L : s_R, s_MP = left ? s_M : s_P;
        double s_S, iden_raw;
        if (!ws2) {
            double iB, iX;
            rcp_nr2(B, fma(s_K, B, -A), iB, iX);
            s_S = A * iB;
            iden_raw = B * iX;
        } else {
            s_S = s_S_ws2;
            iden_raw = rcp_nr(s_K - s_S);
        }
        const double gap = s_K - s_S;  // the reference clamps it away from zero: min(., -eps) / max(., eps)
        const bool clamped = left ? !(gap < -MFC5_SGM_EPS) : !(gap > MFC5_SGM_EPS);
        const double iden = clamped ? (left ? -1.0 / MFC5_SGM_EPS : 1.0 / MFC5_SGM_EPS) : iden_raw;
        const double xi_K = (s_K - u_K) * iden;
        const double xi_K_m1 = (s_S - u_K) * iden;
        const double adv = fma(s_MP, xi_K_m1, u_K);  // u_K + s_MP (xi_K - 1)
#pragma unroll
        for (int f = 0; f < NF; f++) { F[f] = qK[f] * adv; F[IA + f] = qK[IA + f] * adv; }
#pragma unroll
        for (int a = 0; a < ND; a++) {
            if (a == DIR) F[IM + a] = rho_K * (u_K * u_K + s_MP * s_K * xi_K_m1) + p_K;
            else F[IM + a] = rho_K * (qK[IM + a] * adv);
        }
        F[IE] = u_K * (E_K + p_K) + s_MP * (E_K * xi_K_m1 + xi_K * (s_S - u_K) * (rho_K * s_S + p_K * rcp_nr(s_K - u_K)));
        if (EXT && P.low_mach == 1) {   // m_riemann_solver_hllc.fpp:1286-1305 (5-equation emission)
            F[IM + DIR] = F[IM + DIR] + pcw;
            F[IE] = F[IE] + pcw * s_S;
        }
        Fs[0] = adv;
    } else {
        const double inv = rcp_nr(s_M - s_P);
        const double smp = s_M * s_P;
#pragma unroll
        for (int f = 0; f < NF; f++) F[f] = (s_M * arho_R[f] * uR - s_P * arho_L[f] * uL + smp * (arho_L[f] - arho_R[f])) * inv;
#pragma unroll
        for (int a = 0; a < ND; a++) {
            const double flg = (a == DIR) ? 1.0 : 0.0;
            F[IM + a] = (s_M * (rho_R * uR * qR[IM + a] + flg * pres_R) - s_P * (rho_L * uL * qL[IM + a] + flg * pres_L) +
                         smp * (rho_L * qL[IM + a] - rho_R * qR[IM + a])) * inv;
        }
        F[IE] = (s_M * uR * (E_R + pres_R) - s_P * uL * (E_L + pres_L) + smp * (E_L - E_R)) * inv;
        if (EXT && P.low_mach == 1) {   // m_riemann_solver_hll.fpp:446-455,496-515
#pragma unroll
            for (int a = 0; a < ND; a++) F[IM + a] = F[IM + a] + pcw * (qR[IM + a] - qL[IM + a]);
            F[IE] = F[IE] + pcw * (vel_R_rms - vel_L_rms) * 5.e-1;
        }
        if (!P.hll_u) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                F[IA + f] = (qL[IA + f] - qR[IA + f]) * smp * inv;
                Fs[f] = (s_M * qR[IA + f] - s_P * qL[IA + f]) * inv;
            }
        } else {
            const double iw = rcp_nr(s_R - s_L);
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const double aL = qL[IA + f], aR = qR[IA + f];
                if (0.0 <= s_L) F[IA + f] = aL * uL;
                else if (s_R <= 0.0) F[IA + f] = aR * uR;
                else F[IA + f] = (s_R * aL * uL - s_L * aR * uR + s_L * s_R * (aR - aL)) * iw;
            }
            if (0.0 <= s_L) Fs[0] = uL;
            else if (s_R <= 0.0) Fs[0] = uR;
            else Fs[0] = (s_R * uL - s_L * uR) * iw;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// One Riemann problem on a face normal to DIR.
// qL = right-boundary value of the left cell, qR = left-boundary value of the right cell.
// HLLC: src/simulation/m_riemann_solver_hllc.fpp:878-1332 (5-eq pure-fluid emission)
// HLL : src/simulation/m_riemann_solver_hll.fpp:128-576
// averages: src/simulation/include/inline_riemann.fpp:1-30
// F[0..N) fluxes; Fs[0] face-normal velocity (u-interface modes) or Fs[0..NF) interface alphas (HLL method 1)
// -------------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------------
// Viscous stresses (viscous = T, weno_Re_flux = F; reference-order kernels only).
// What the Riemann solver exports for the viscous source flux of its face: the interface velocities vel_src_rsx_vf
// (hllc.fpp:1317-1322; the HLL solver never writes them: zero) and the interface Reynolds numbers Re_avg_rsx_vf
// (hllc.fpp:975-977,1073-1080; hll.fpp:205-207,340-347).
// -------------------------------------------------------------------------------------------------
struct VisFace {
    double vel_src[3];
    double re_avg[2];
};

// s_compute_interface_reynolds, m_riemann_state.fpp:1062-1090
template <int NF>
__device__ __forceinline__ void interface_reynolds(const DevP& P, const double* alpha, double* Re) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        Re[i] = -1.e6;  // dflt_real
        if (P.re_size[i] > 0) Re[i] = 0.0;
        for (int q = 0; q < P.re_size[i]; q++) {
            double a = 0.0;
#pragma unroll
            for (int f = 0; f < NF; f++) if (f == P.re_idx[i][q]) a = alpha[f];
            Re[i] = a / P.res[i][q] + Re[i];
        }
        Re[i] = 1.0 / fmax(Re[i], MFC5_SGM_EPS);
    }
}

// Viscous source flux of the momentum and energy equations on the face between cell c (coordinates idx[]) and its neighbour
// c + e_DIR: s_compute_cartesian_viscous_source_flux m_riemann_state.fpp:805-975 with s_calculate_shear / bulk_stress_tensor
// (:978-1031), fed by the finite-difference cell-boundary velocity gradients of s_get_viscous (m_viscous.fpp:515-830).
// The solver averages the RIGHT-boundary gradient field of cell c and the LEFT-boundary one of cell c + e_DIR (argument swap,
// m_rhs.fpp:1031-1033); both are the same expression of the cell-centre velocities, evaluated here from the primitive
// variables themselves (read through L1 / L2):
//   along DIR      (v(c + e) - v(c)) / (cc(j + 1) - cc(j))
//   along dd != DIR 0.25 (L(c + e) + R(c + e) + L(c) + R(c)),  L(x) = (v(x) - v(x - e_dd)) / (cc - cc'),  R(x) = (v(x + e_dd) - v(x)) / (cc' - cc)
// Fv[0 .. ND-1]: momentum components, Fv[ND]: energy.
template <int NF, int ND, int DIR>
__device__ __forceinline__ void viscous_face(const DevP& P, long long c, const int* idx, const VisFace& vo, double* Fv) {
    constexpr int IM = NF;
    const long long strd[3] = {1, P.sy, P.sz};
    const long long sd = strd[DIR];
    double vga[ND][ND];
#pragma unroll
    for (int a = 0; a < ND; a++) {
        const double* v = P.prim + (long long)(IM + a) * P.fs;
#pragma unroll
        for (int dd = 0; dd < ND; dd++) {
            const double* cc = P.T[dd].cc + P.b;
            const int jj = idx[dd];
            double g;
            if (dd == DIR) {
                g = (__ldg(v + c + sd) - __ldg(v + c)) / (cc[jj + 1] - cc[jj]);
            } else {
                const long long st = strd[dd];
                const double hL = cc[jj] - cc[jj - 1], hR = cc[jj + 1] - cc[jj];
                const double L1 = (__ldg(v + c + sd) - __ldg(v + c + sd - st)) / hL, R1 = (__ldg(v + c + sd + st) - __ldg(v + c + sd)) / hR;
                const double L0 = (__ldg(v + c) - __ldg(v + c - st)) / hL, R0 = (__ldg(v + c + st) - __ldg(v + c)) / hR;
                g = 25.e-2 * (L1 + R1 + L0 + R0);
            }
            vga[a][dd] = 0.5 * (g + g);
        }
    }
    double div = 0.0;
#pragma unroll
    for (int a = 0; a < ND; a++) div = div + vga[a][a];
#pragma unroll
    for (int a = 0; a <= ND; a++) Fv[a] = 0.0;   // s_initialize_riemann_solver, m_riemann_state.fpp:440-450
    if (P.re_size[0] > 0) {   // shear_stress
#pragma unroll
        for (int a = 0; a < ND; a++) {
            double tau = (vga[a][DIR] + vga[DIR][a]) / vo.re_avg[0];
            if (a == DIR) tau = tau - (2.0 / 3.0) * div / vo.re_avg[0];
            Fv[a] = Fv[a] - tau;
            Fv[ND] = Fv[ND] - vo.vel_src[a] * tau;
        }
    }
    if (P.re_size[1] > 0) {   // bulk_stress
#pragma unroll
        for (int a = 0; a < ND; a++) {
            const double tau = (a == DIR) ? div / vo.re_avg[1] : 0.0;
            Fv[a] = Fv[a] - tau;
            Fv[ND] = Fv[ND] - vo.vel_src[a] * tau;
        }
    }
}

template <int NF, int ND, int DIR, int RS>
__device__ __forceinline__ void riemann_face(const DevP& P, double* qL, double* qR, double* F, double* Fs, VisFace* vo = nullptr) {
    constexpr int IM = NF, IE = NF + ND, IA = NF + ND + 1;
    const int rs = (RS < 0) ? P.rs : RS;
    double vel_L_rms = 0.0, vel_R_rms = 0.0;
#pragma unroll
    for (int a = 0; a < ND; a++) {
        vel_L_rms = vel_L_rms + qL[IM + a] * qL[IM + a];
        vel_R_rms = vel_R_rms + qR[IM + a] * qR[IM + a];
    }
    const double pres_L = qL[IE], pres_R = qR[IE];
    double arho_L[NF], arho_R[NF], alf_L[NF], alf_R[NF];
    double adv_L[NF], adv_R[NF];  // alpha_L/R handed to s_compute_speed_of_sound: hllc.fpp:889-892 loads them before the limiter
#pragma unroll
    for (int f = 0; f < NF; f++) { adv_L[f] = qL[IA + f]; adv_R[f] = qR[IA + f]; }
    if (rs == 2) {
        if (P.mpp) {  // in-place limiter on the face states, hllc.fpp:938-958
            double sL = 0.0, sR = 0.0;
#pragma unroll
            for (int f = 0; f < NF; f++) {
                qL[f] = fmax(0.0, qL[f]);
                qL[IA + f] = fmin(fmax(0.0, qL[IA + f]), 1.0);
                qR[f] = fmax(0.0, qR[f]);
                qR[IA + f] = fmin(fmax(0.0, qR[IA + f]), 1.0);
                sL = sL + qL[IA + f];
                sR = sR + qR[IA + f];
            }
#pragma unroll
            for (int f = 0; f < NF; f++) {
                qL[IA + f] = qL[IA + f] / fmax(sL, MFC5_SGM_EPS);
                qR[IA + f] = qR[IA + f] / fmax(sR, MFC5_SGM_EPS);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; f++) {
            arho_L[f] = qL[f]; arho_R[f] = qR[f];
            alf_L[f] = qL[IA + f]; alf_R[f] = qR[IA + f];
        }
    } else {
#pragma unroll
        for (int f = 0; f < NF; f++) {
            arho_L[f] = qL[f]; arho_R[f] = qR[f];
            alf_L[f] = qL[IA + f]; alf_R[f] = qR[IA + f];
        }
        if (P.mpp) {  // HLL limits local copies only, hll.fpp:187-200
            double sL = 0.0, sR = 0.0;
#pragma unroll
            for (int f = 0; f < NF; f++) {
                arho_L[f] = fmax(0.0, arho_L[f]);
                alf_L[f] = fmin(fmax(0.0, alf_L[f]), 1.0);
                sL = sL + alf_L[f];
                arho_R[f] = fmax(0.0, arho_R[f]);
                alf_R[f] = fmin(fmax(0.0, alf_R[f]), 1.0);
                sR = sR + alf_R[f];
            }
#pragma unroll
            for (int f = 0; f < NF; f++) {
                alf_L[f] = alf_L[f] / fmax(sL, MFC5_SGM_EPS);
                alf_R[f] = alf_R[f] / fmax(sR, MFC5_SGM_EPS);
            }
        }
    }
    if (rs != 2) {  // hll.fpp:187-200 limits alpha_L/R themselves
#pragma unroll
        for (int f = 0; f < NF; f++) { adv_L[f] = alf_L[f]; adv_R[f] = alf_R[f]; }
    }
    // s_accumulate_mixture_properties, m_riemann_state.fpp:1036-1058
    double rho_L = 0, gamma_L = 0, pi_inf_L = 0, qv_L = 0, rho_R = 0, gamma_R = 0, pi_inf_R = 0, qv_R = 0;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        rho_L = rho_L + arho_L[f];
        gamma_L = gamma_L + alf_L[f] * P.gam[f];
        pi_inf_L = pi_inf_L + alf_L[f] * P.pinf[f];
        qv_L = qv_L + arho_L[f] * P.qv[f];
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        rho_R = rho_R + arho_R[f];
        gamma_R = gamma_R + alf_R[f] * P.gam[f];
        pi_inf_R = pi_inf_R + alf_R[f] * P.pinf[f];
        qv_R = qv_R + arho_R[f] * P.qv[f];
    }
    const double E_L = gamma_L * pres_L + pi_inf_L + 5.e-1 * rho_L * vel_L_rms + qv_L;
    const double E_R = gamma_R * pres_R + pi_inf_R + 5.e-1 * rho_R * vel_R_rms + qv_R;
    const double H_L = (E_L + pres_L) / rho_L;
    const double H_R = (E_R + pres_R) / rho_R;
    const double c_L = sound_speed(P, pres_L, rho_L, gamma_L, H_L, adv_L, vel_L_rms, qv_L);
    const double c_R = sound_speed(P, pres_R, rho_R, gamma_R, H_R, adv_R, vel_R_rms, qv_R);
    if (vo) {
        double Re_L[2], Re_R[2];
        interface_reynolds<NF>(P, adv_L, Re_L);
        interface_reynolds<NF>(P, adv_R, Re_R);
#pragma unroll
        for (int q = 0; q < 2; q++) vo->re_avg[q] = 2.0 / (1.0 / Re_L[q] + 1.0 / Re_R[q]);
    }
    // low Mach corrections, inline_riemann.fpp:71-93 (hllc.fpp:1084-1086,1137-1141; hll.fpp:400-404)
    double zcoef = 0.0;
    if (P.low_mach != 0) zcoef = fmin(1.0, fmax(sqrt(vel_L_rms) / c_L, sqrt(vel_R_rms) / c_R));
    if (rs == 2 && P.low_mach == 2) {  // the face-normal velocities of the local copies are replaced
        const double vel_L_tmp = 5.e-1 * ((qL[IM + DIR] + qR[IM + DIR]) + zcoef * (qL[IM + DIR] - qR[IM + DIR]));
        const double vel_R_tmp = 5.e-1 * ((qL[IM + DIR] + qR[IM + DIR]) + zcoef * (qR[IM + DIR] - qL[IM + DIR]));
        qL[IM + DIR] = vel_L_tmp;
        qR[IM + DIR] = vel_R_tmp;
    }
    const double uL = qL[IM + DIR], uR = qR[IM + DIR];
    double s_L, s_R, s_S;
    if (P.ws == 1) {
        s_L = fmin(uL - c_L, uR - c_R);
        s_R = fmax(uR + c_R, uL + c_L);
        s_S = (pres_R - pres_L + rho_L * uL * (s_L - uL) - rho_R * uR * (s_R - uR)) / (rho_L * (s_L - uL) - rho_R * (s_R - uR));
    } else {
        double rho_avg, vel_avg_rms, H_avg, gamma_avg, qv_avg;
        if (P.avg == 1) {  // Roe
            const double sl = sqrt(rho_L), sr = sqrt(rho_R);
            rho_avg = sqrt(rho_L * rho_R);
            H_avg = (sl * H_L + sr * H_R) / (sl + sr);
            gamma_avg = (sl * gamma_L + sr * gamma_R) / (sl + sr);
            vel_avg_rms = sq(sl * qL[IM] + sr * qR[IM]) / sq(sl + sr);  // component 1 only, inline_riemann.fpp:28
            qv_avg = (sl * qv_L + sr * qv_R) / (sl + sr);
        } else {
            rho_avg = 5.e-1 * (rho_L + rho_R);
            vel_avg_rms = 0.0;
#pragma unroll
            for (int a = 0; a < ND; a++) vel_avg_rms = vel_avg_rms + sq(5.e-1 * (qL[IM + a] + qR[IM + a]));
            H_avg = 5.e-1 * (H_L + H_R);
            gamma_avg = 5.e-1 * (gamma_L + gamma_R);
            qv_avg = 5.e-1 * (qv_L + qv_R);
        }
        const double c_avg = sound_speed(P, pres_R, rho_avg, gamma_avg, H_avg, adv_R, vel_avg_rms, qv_avg);  // pres_R, alpha_R: placeholders, hllc.fpp:1063-1066
        const double pres_SL = 5.e-1 * (pres_L + pres_R + rho_avg * c_avg * (uL - uR));
        const double pres_SR = pres_SL;
        const double Ms_L = fmax(1.0, sqrt(1.0 + ((5.e-1 + gamma_L) / (1.0 + gamma_L)) * (pres_SL / pres_L - 1.0) * pres_L /
                                                     ((pres_L + pi_inf_L / (1.0 + gamma_L)))));
        const double Ms_R = fmax(1.0, sqrt(1.0 + ((5.e-1 + gamma_R) / (1.0 + gamma_R)) * (pres_SR / pres_R - 1.0) * pres_R /
                                                     ((pres_R + pi_inf_R / (1.0 + gamma_R)))));
        s_L = uL - c_L * Ms_L;
        s_R = uR + c_R * Ms_R;
        s_S = 5.e-1 * ((uL + uR) + (pres_L - pres_R) / (rho_avg * c_avg));
    }
    double s_M = fmin(0.0, s_L), s_P = fmax(0.0, s_R);
    if (rs == 5) {  // Lax-Friedrichs: one symmetric speed, m_riemann_solver_lf.fpp:238-253
        s_P = fmax(sqrt(vel_L_rms), sqrt(vel_R_rms)) + fmax(c_L, c_R);
        s_M = -s_P;
        s_L = s_M; s_R = s_P;
    }
    double pcorr = 0.0;
    if (P.low_mach == 1) {
        if (rs == 2) pcorr = rho_L * rho_R * (s_L - uL) * (s_R - uR) * (uR - uL) / (rho_R * (s_R - uR) - rho_L * (s_L - uL)) * (zcoef - 1.0);
        else pcorr = -(s_P - s_M) * (rho_L + rho_R) / 8.0 * (zcoef - 1.0);
    }
    if (rs == 2) {
        const double xi_L = (s_L - uL) / fmin(s_L - s_S, -MFC5_SGM_EPS);
        const double xi_R = (s_R - uR) / fmax(s_R - s_S, MFC5_SGM_EPS);
        const double xi_L_m1 = (s_S - uL) / fmin(s_L - s_S, -MFC5_SGM_EPS);
        const double xi_R_m1 = (s_S - uR) / fmax(s_R - s_S, MFC5_SGM_EPS);
        const double xi_M = (5.e-1 + fsign(5.e-1, s_S));
        const double xi_P = (5.e-1 - fsign(5.e-1, s_S));
#pragma unroll
        for (int f = 0; f < NF; f++) F[f] = xi_M * qL[f] * (uL + s_M * xi_L_m1) + xi_P * qR[f] * (uR + s_P * xi_R_m1);
#pragma unroll
        for (int a = 0; a < ND; a++) {
            if (a == DIR)
                F[IM + a] = xi_M * (rho_L * (uL * qL[IM + a] + s_M * (s_L)*xi_L_m1) + (pres_L)) +
                            xi_P * (rho_R * (uR * qR[IM + a] + s_P * (s_R)*xi_R_m1) + (pres_R));
            else
                F[IM + a] = xi_M * (rho_L * (uL * qL[IM + a] + s_M * (qL[IM + a]) * xi_L_m1)) +
                            xi_P * (rho_R * (uR * qR[IM + a] + s_P * (qR[IM + a]) * xi_R_m1));
        }
        F[IE] = xi_M * (uL * (E_L + pres_L) + s_M * (E_L * xi_L_m1 + xi_L * (s_S - uL) * (rho_L * s_S + pres_L / (s_L - uL)))) +
                xi_P * (uR * (E_R + pres_R) + s_P * (E_R * xi_R_m1 + xi_R * (s_S - uR) * (rho_R * s_S + pres_R / (s_R - uR))));
        if (P.low_mach == 1) {
            F[IM + DIR] = F[IM + DIR] + (s_M / s_L) * (s_P / s_R) * pcorr;
            F[IE] = F[IE] + (s_M / s_L) * (s_P / s_R) * pcorr * s_S;
        }
#pragma unroll
        for (int f = 0; f < NF; f++)
            F[IA + f] = xi_M * qL[IA + f] * (uL + s_M * xi_L_m1) + xi_P * qR[IA + f] * (uR + s_P * xi_R_m1);
        Fs[0] = xi_M * (uL + s_M * xi_L_m1) + xi_P * (uR + s_P * xi_R_m1);
        if (vo) {
#pragma unroll
            for (int a = 0; a < ND; a++) {
                const double flg = (a == DIR) ? 1.0 : 0.0;
                vo->vel_src[a] = xi_M * (qL[IM + a] + flg * s_M * xi_L_m1) + xi_P * (qR[IM + a] + flg * s_P * xi_R_m1);
            }
        }
    } else {
        if (vo) {
#pragma unroll
            for (int a = 0; a < ND; a++) vo->vel_src[a] = 0.0;
        }
#pragma unroll
        for (int f = 0; f < NF; f++)
            F[f] = (s_M * arho_R[f] * uR - s_P * arho_L[f] * uL + s_M * s_P * (arho_L[f] - arho_R[f])) / (s_M - s_P);
#pragma unroll
        for (int a = 0; a < ND; a++) {
            const double flg = (a == DIR) ? 1.0 : 0.0;
            F[IM + a] = (s_M * (rho_R * uR * qR[IM + a] + flg * pres_R) - s_P * (rho_L * uL * qL[IM + a] + flg * pres_L) +
                         s_M * s_P * (rho_L * qL[IM + a] - rho_R * qR[IM + a])) / (s_M - s_P);
        }
        F[IE] = (s_M * uR * (E_R + pres_R) - s_P * uL * (E_L + pres_L) + s_M * s_P * (E_L - E_R)) / (s_M - s_P);
        if (P.low_mach == 1) {
#pragma unroll
            for (int a = 0; a < ND; a++) F[IM + a] = F[IM + a] + (s_M / s_L) * (s_P / s_R) * pcorr * (qR[IM + a] - qL[IM + a]);
            F[IE] = F[IE] + (s_M / s_L) * (s_P / s_R) * pcorr * (vel_R_rms - vel_L_rms) / 2.0;
        }
        // Lax-Friedrichs always exports the interface alphas (m_riemann_solver_lf.fpp:306-313); the RHS then reads
        // flux_src(adv%beg) as the shared face velocity (adv_src_mode_vel_iface, m_global_parameters.fpp:903-906)
        if (!P.hll_u || rs == 5) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                F[IA + f] = (qL[IA + f] - qR[IA + f]) * s_M * s_P / (s_M - s_P);
                Fs[f] = (s_M * qR[IA + f] - s_P * qL[IA + f]) / (s_M - s_P);
            }
        } else {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const double aL = qL[IA + f], aR = qR[IA + f];
                if (0.0 <= s_L) F[IA + f] = aL * uL;
                else if (s_R <= 0.0) F[IA + f] = aR * uR;
                else F[IA + f] = (s_R * aL * uL - s_L * aR * uR + s_L * s_R * (aR - aL)) / (s_R - s_L);
            }
            if (0.0 <= s_L) Fs[0] = uL;
            else if (s_R <= 0.0) Fs[0] = uR;
            else Fs[0] = (s_R * uL - s_L * uR) / (s_R - s_L);
        }
    }
}

}  // namespace mfc5
