// mfc5_kernels.cuh -- the kernels of one RK stage (sm_100a, FP64):
//   k_cons2prim   cons -> prim on the interior                         (reference K1+K2)
//   k_bc_fill     ghost-cell fill of prims for one (direction, side)   (K3)
//   k_cbc         characteristic-BC replacement fluxes of one boundary plane (K9)
//   k_sweep_x     fused WENO + Riemann + flux difference + alpha source [+ RK] along x: one warp
//                 owns 30 consecutive cells, neighbour exchange by warp shuffles       (K4..K8,K10,K11[,K12])
//   k_sweep_march same along any direction: each thread marches a pencil; the stencil window of the
//                 primitive variables is staged in a shared-memory ring, next plane prefetched
//   k_reduce      ICFL max / dt min / NaN count                         (K13)
// Face states and face fluxes never touch HBM.
#pragma once
#include "mfc5_dev.cuh"

namespace mfc5 {

// index type of the sweep kernels' field accesses (see k_sweep_march)
template <bool NARROW> struct SweepIdx { using type = long long; };
template <> struct SweepIdx<true> { using type = unsigned; };

template <int NF, int ND>
struct Shape {
    static constexpr int N = 2 * NF + ND + 1;
    static constexpr int IM = NF, IE = NF + ND, IA = NF + ND + 1;
};

// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, completion by group
__device__ __forceinline__ void cp_async8(double* sdst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* sdst, const double* gsrc) {   // two consecutive doubles, both addresses 16-byte aligned
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#ifdef MFC5_NO_PREFETCH
__device__ __forceinline__ void prefetch_l2(const double*) {}
#else
__device__ __forceinline__ void prefetch_l2(const double* g) { asm volatile("prefetch.global.L2 [%0];" ::"l"(g)); }
#endif

// mixture rule: s_compute_species_fraction (m_variables_conversion.fpp:1140-1188) followed by
// s_convert_species_to_mixture_variables_kernel (:198-270); `passes` = number of limiter passes.
template <int NF>
__device__ __forceinline__ void mixture(const DevP& P, double* arho, double* alf, int passes, double& rho, double& gam, double& pinf,
                                        double& qv) {
    if (P.mpp) {
        for (int pass = 0; pass < passes; pass++) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < NF; i++) {
                arho[i] = fmax(0.0, arho[i]);
                alf[i] = fmin(fmax(0.0, alf[i]), 1.0);
                sum = sum + alf[i];
            }
#pragma unroll
            for (int i = 0; i < NF; i++) alf[i] = alf[i] / fmax(sum, MFC5_SGM_EPS);
        }
    }
    rho = 0; gam = 0; pinf = 0; qv = 0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        rho = rho + arho[i];
        gam = gam + alf[i] * P.gam[i];
        pinf = pinf + alf[i] * P.pinf[i];
        qv = qv + arho[i] * P.qv[i];
    }
}

// -------------------------------------------------------------------------------------------------
// cons -> prim, interior cells: s_convert_conservative_to_primitive_variables,
// src/common/m_variables_conversion.fpp:459-717 (density floor on: m_start_up.fpp:823)
// -------------------------------------------------------------------------------------------------
// one cell: q[N] conservative -> pr[N] primitive (the stored partial densities / volume fractions are the unclamped ones)
template <int NF, int ND>
__device__ __forceinline__ void cons2prim_cell(const DevP& P, const double* q, double* pr) {
    using S = Shape<NF, ND>;
    double arho[NF], alf[NF], rho, gam, pinf, qv;
#pragma unroll
    for (int f = 0; f < NF; f++) { arho[f] = q[f]; alf[f] = q[S::IA + f]; }
    mixture<NF>(P, arho, alf, 2, rho, gam, pinf, qv);
    rho = fmax(rho, MFC5_SGM_EPS);
    double dyn = 0.0;
#pragma unroll
    for (int a = 0; a < ND; a++) {
        const double u = q[S::IM + a] / rho;
        dyn = dyn + 5.e-1 * q[S::IM + a] * u;
        pr[S::IM + a] = u;
    }
    pr[S::IE] = (q[S::IE] - dyn - pinf - qv) / gam;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        pr[f] = q[f];
        pr[S::IA + f] = q[S::IA + f];
    }
}

template <int NF, int ND>
__global__ void __launch_bounds__(256) k_cons2prim(const __grid_constant__ DevP P) {
    using S = Shape<NF, ND>;
    const int nx = P.s[0] + 1, ny = P.s[1] + 1, nz = P.s[2] + 1;
    const long long tot = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx);
        const long long r = t / nx;
        const int j = (int)(r % ny), k = (int)(r / ny);
        const long long c = P.base + i + j * P.sy + k * P.sz;
        double q[S::N], pr[S::N];
#pragma unroll
        for (int v = 0; v < S::N; v++) q[v] = P.q1[v * P.fs + c];
        cons2prim_cell<NF, ND>(P, q, pr);
#pragma unroll
        for (int v = 0; v < S::N; v++) P.primw[v * P.fs + c] = pr[v];
    }
}

// -------------------------------------------------------------------------------------------------
// Ghost-cell fill for one (direction, side): s_populate_bc_direction m_boundary_common.fpp:120-187,
// kernels of m_boundary_primitives.fpp (:27 extrapolation, :114 symmetry, :338 periodic, :552 slip,
// :704 no-slip, :877 Dirichlet).  One thread per (transverse point, variable); loops the b layers.
// -------------------------------------------------------------------------------------------------
#ifdef MFC5_HOST_TU
// `full` = 1: the reference's sequential semantics (direction d spans the ghosts of directions < d, so
// edges and corners get filled; needed only when a caller asks for the ghosted q_prim_vf);
// `full` = 0: transverse interior only -- all the inviscid sweeps, CBC and the reductions ever read
// (faces use the transverse interior, m_rhs.fpp:1020-1027), which makes the six faces independent.
__device__ __forceinline__ void bc_fill_face(const DevP& P, int d, int side, int full, long long t0, long long nthreads) {
    const int bc = P.bc[d][side];
    int lo[3] = {0, 0, 0}, hi[3] = {P.s[0], P.s[1], P.s[2]};
    if (full) for (int e = 0; e < d; e++) { lo[e] = -P.b; hi[e] = P.s[e] + P.b; }
    const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    const int n1 = hi[t1] - lo[t1] + 1, n2 = hi[t2] - lo[t2] + 1;
    const long long tot = (long long)n1 * n2 * P.N;
    const long long strd[3] = {1, P.sy, P.sz};
    const int s = P.s[d];
    for (long long t = t0; t < tot; t += nthreads) {
        const int a1 = (int)(t % n1) + lo[t1];
        const long long r = t / n1;
        const int a2 = (int)(r % n2) + lo[t2];
        const int v = (int)(r / n2);
        double* f = P.primw + v * P.fs + P.base + a1 * strd[t1] + a2 * strd[t2];
        const int comp = v - P.nf;  // velocity component if in [0, nd)
        const double wv_n = side == 0 ? P.vb[d][d] : P.ve[d][d];
        for (int l = 1; l <= P.b; l++) {
            const int g = side == 0 ? -l : s + l;
            const int edge = side == 0 ? 0 : s;
            const int mir = side == 0 ? l - 1 : s - (l - 1);
            const int per = side == 0 ? s - (l - 1) : l - 1;
            double val;
            switch (bc) {
            case MFC5_BC_PERIODIC: val = f[per * strd[d]]; break;
            case MFC5_BC_REFLECTIVE: val = (comp == d) ? -f[mir * strd[d]] : f[mir * strd[d]]; break;
            case MFC5_BC_SLIP_WALL: val = (comp == d) ? -f[mir * strd[d]] + 2.0 * wv_n : f[edge * strd[d]]; break;
            case MFC5_BC_NO_SLIP_WALL:
                if (comp >= 0 && comp < P.nd) val = -f[mir * strd[d]] + 2.0 * (side == 0 ? P.vb[d][comp] : P.ve[d][comp]);
                else val = f[edge * strd[d]];
                break;
            case MFC5_BC_DIRICHLET:
                if (P.dirich[d][side]) {
                    const int m1 = P.s[t1] + 1, m2 = P.s[t2] + 1;
                    if (a1 < 0 || a1 >= m1 || a2 < 0 || a2 >= m2) val = 0.0;
                    else val = P.dirich[d][side][((long long)v * m2 + a2) * m1 + a1];
                } else val = f[edge * strd[d]];
                break;
            default: val = f[edge * strd[d]]; break;
            }
            f[g * strd[d]] = val;
        }
    }
}

// one (direction, side), reference order (used for the ghosted q_prim_vf download)
__global__ void __launch_bounds__(256) k_bc_fill(const __grid_constant__ DevP P, int d, int side, int full) {
    bc_fill_face(P, d, side, full, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// the physical faces face0 .. face0 + gridDim.y - 1 of the block in ONE launch: face = 2 * direction + side
__global__ void __launch_bounds__(256) k_bc_faces(const __grid_constant__ DevP P, int face0) {
    const int face = blockIdx.y + face0;
    const int d = face >> 1, side = face & 1;
    if (d >= P.nd || P.bc[d][side] >= 0 || P.gmode[d][side] != MFC5_GM_MEM) return;   // folded faces: the sweep loaders do it
    bc_fill_face(P, d, side, 0, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

#endif  // MFC5_HOST_TU

// -------------------------------------------------------------------------------------------------
// Characteristic boundary conditions: s_cbc src/simulation/m_cbc.fpp:464-927 (+ s_initialize_cbc
// :930-1182 sign conventions, s_finalize_cbc :1185-1345), cell-centre fluxes
// m_variables_conversion.fpp:952-1135, wave amplitudes m_compute_cbc.fpp:136-365.
// For WENO3/5 the rewritten faces depend on the primitive variables only, so this kernel runs BEFORE
// the sweep of direction d and stores the replacement flux / flux_src of local faces -1..nface-1 in
// cbcbuf; the sweep kernel substitutes them for its Riemann fluxes.  One thread per plane point.
// -------------------------------------------------------------------------------------------------
struct CbcCoef {
    double fd[8];
    double pi[2][3];
    double ds0, Del;
    double vel_in[3], vel_out[3], pres_in, pres_out, arho_in[MFC5_MAX_FLUIDS], alf_in[MFC5_MAX_FLUIDS];
    int grcbc_in, grcbc_out, grcbc_vel_out;
};

// one plane point of the face (direction d, compile-time: every velocity-component index below is a constant, so the layer and
// flux arrays stay in registers)
template <int NF, int ND, int d>
__device__ __forceinline__ void cbc_point(const DevP& P, const CbcCoef& K, const int side) {
    using S = Shape<NF, ND>;
    constexpr int N = S::N;
    constexpr int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    const int n1 = P.s[t1] + 1, n2 = P.s[t2] + 1;
    const long long plane = (long long)n1 * n2;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= plane) return;
    const int a1 = (int)(t % n1), a2 = (int)(t / n1);
    const long long strd[3] = {1, P.sy, P.sz};
    const long long base = P.base + a1 * strd[t1] + a2 * strd[t2];
    const long long sd = strd[d];
    const int s = P.s[d], b = P.b, bc = P.bc[d][side];
    const double sgn = side == 0 ? 1.0 : -1.0;
    const bool hll1 = (P.rs == 1 && !P.hll_u);
    const int nface = P.nface;
    constexpr int didx[3] = {d, d == 0 ? 1 : 0, d == 2 ? 1 : 2};

    double q[7][N], Fc[4][N], Fsc[4][NF];   // layers 0 .. buff_size (<= 6: viscous WENO5)
#pragma unroll
    for (int jl = 0; jl < 7; jl++) {
        if (jl <= b) {
            const long long c = base + sd * (side == 0 ? jl : s - jl);
#pragma unroll
            for (int v = 0; v < N; v++) q[jl][v] = P.prim[v * P.fs + c];
            q[jl][S::IM + d] = q[jl][S::IM + d] * sgn;
        } else {
#pragma unroll
            for (int v = 0; v < N; v++) q[jl][v] = 0.0;
        }
    }
#pragma unroll
    for (int jl = 0; jl <= 3; jl++) {
        if (jl > b) break;
        double arho[NF], alf[NF], vel[3] = {0, 0, 0}, vsum = 0, rho, gam, pinf, qv;
        for (int f = 0; f < NF; f++) { arho[f] = q[jl][f]; alf[f] = q[jl][S::IA + f]; }
        for (int a = 0; a < ND; a++) vel[a] = q[jl][S::IM + a];
        for (int a = 0; a < ND; a++) vsum = vsum + vel[a] * vel[a];
        const double pres = q[jl][S::IE];
        mixture<NF>(P, arho, alf, 1, rho, gam, pinf, qv);
        const double E = gam * pres + pinf + 5.e-1 * rho * vsum + qv;
        for (int f = 0; f < NF; f++) Fc[jl][f] = arho[f] * vel[d];
        for (int a = 0; a < ND; a++) Fc[jl][S::IM + a] = rho * vel[d] * vel[a] + pres * (a == d ? 1.0 : 0.0);
        Fc[jl][S::IE] = vel[d] * (E + pres);
        if (hll1) {
            for (int f = 0; f < NF; f++) { Fc[jl][S::IA + f] = 0.0; Fsc[jl][f] = alf[f]; }
        } else {
            for (int f = 0; f < NF; f++) { Fc[jl][S::IA + f] = vel[d] * alf[f]; Fsc[jl][f] = vel[d]; }
        }
    }
    double fl[3][N], fsl[3][NF];  // local faces -1, 0, 1 -> index +1
    if (nface == 1) {
        for (int v = 0; v < N; v++) fl[1][v] = Fc[0][v] + K.pi[0][0] * (Fc[1][v] - Fc[0][v]);
        for (int f = 0; f < NF; f++) fsl[1][f] = Fsc[0][f] + (Fsc[1][f] - Fsc[0][f]) * K.pi[0][0];
    } else {
        for (int jl = 0; jl <= 1; jl++) {
            for (int v = 0; v < N; v++)
                fl[jl + 1][v] = Fc[jl][v] + K.pi[jl][0] * (Fc[3][v] - Fc[2][v]) + K.pi[jl][1] * (Fc[2][v] - Fc[1][v]) +
                                K.pi[jl][2] * (Fc[1][v] - Fc[0][v]);
            for (int f = 0; f < NF; f++)
                fsl[jl + 1][f] = Fsc[jl][f] + (Fsc[3][f] - Fsc[2][f]) * K.pi[jl][0] + (Fsc[2][f] - Fsc[1][f]) * K.pi[jl][1] +
                                 (Fsc[1][f] - Fsc[0][f]) * K.pi[jl][2];
        }
    }
    double arho[NF], adv[NF], vel[3] = {0, 0, 0}, vel_K_sum = 0, rho, gam, pinf, qv;
    for (int f = 0; f < NF; f++) arho[f] = q[0][f];
    for (int a = 0; a < ND; a++) vel[a] = q[0][S::IM + a];
    for (int a = 0; a < ND; a++) vel_K_sum = vel_K_sum + vel[a] * vel[a];
    const double pres = q[0][S::IE];
    for (int f = 0; f < NF; f++) adv[f] = q[0][S::IA + f];
    mixture<NF>(P, arho, adv, 1, rho, gam, pinf, qv);
    double mf[NF];
    for (int f = 0; f < NF; f++) mf[f] = arho[f] / rho;
    const double E = gam * pres + pinf + 5.e-1 * rho * vel_K_sum;
    const double H = (E + pres) / rho;
    const double c = sound_speed(P, pres, rho, gam, H, adv, vel_K_sum, qv);
    double darho_ds[NF], dvel_ds[3] = {0, 0, 0}, dpres_ds = 0, dadv_ds[NF];
    for (int f = 0; f < NF; f++) { darho_ds[f] = 0; dadv_ds[f] = 0; }
#pragma unroll
    for (int jl = 0; jl < 7; jl++) {
        if (jl > b) break;
        for (int f = 0; f < NF; f++) darho_ds[f] = q[jl][f] * K.fd[jl] + darho_ds[f];
        for (int a = 0; a < ND; a++) dvel_ds[a] = q[jl][S::IM + a] * K.fd[jl] + dvel_ds[a];
        dpres_ds = q[jl][S::IE] * K.fd[jl] + dpres_ds;
        for (int f = 0; f < NF; f++) dadv_ds[f] = q[jl][S::IA + f] * K.fd[jl] + dadv_ds[f];
    }
    const double lam1 = vel[d] - c, lam2 = vel[d], lam3 = vel[d] + c;
    const double Ma = vel[d] / c;
    double L[N];
    for (int v = 0; v < N; v++) L[v] = 0.0;
    constexpr int last = N - 1;
    const double L1base = lam1 * (dpres_ds - rho * c * dvel_ds[d]);
    auto fill = [&](double fac, double l2) {
        for (int f = 0; f < NF; f++) L[1 + f] = fac * l2 * (c * c * darho_ds[f] - mf[f] * dpres_ds);
        for (int ii = 1; ii < ND; ii++) L[S::IM + ii] = fac * l2 * dvel_ds[didx[ii]];
        for (int f = 0; f < NF; f++) L[S::IE + f] = fac * l2 * dadv_ds[f];
    };
    switch (bc) {
    case MFC5_BC_CHAR_SLIP_WALL: L[0] = L1base; L[last] = L[0]; break;
    case MFC5_BC_CHAR_NR_SUB_BUFFER: {
        double lf = (5.e-1 - 5.e-1 * fsign(1.0, lam1));
        L[0] = lf * lam1 * (dpres_ds - rho * c * dvel_ds[d]);
        lf = (5.e-1 - 5.e-1 * fsign(1.0, lam2));
        fill(lf, lam2);
        lf = (5.e-1 - 5.e-1 * fsign(1.0, lam3));
        L[last] = lf * lam3 * (dpres_ds + rho * c * dvel_ds[d]);
    } break;
    case MFC5_BC_CHAR_NR_SUB_INFLOW:
        L[0] = L1base;
        if (K.grcbc_in) {
            for (int f = 0; f < NF; f++)
                L[1 + f] = pow(c, 3.0) * Ma * (arho[f] - K.arho_in[f]) / K.Del - c * Ma * (pres - K.pres_in) / K.Del;
            if (ND > 1) {
                L[S::IM + 1] = c * Ma * (vel[didx[1]] - K.vel_in[didx[1]]) / K.Del;
                if (ND > 2) L[S::IM + 2] = c * Ma * (vel[didx[2]] - K.vel_in[didx[2]]) / K.Del;
            }
            for (int f = 0; f < NF; f++) L[S::IE + f] = c * Ma * (adv[f] - K.alf_in[f]) / K.Del;
            L[last] = rho * sq(c) * (1.0 + Ma) * (vel[d] + K.vel_in[d] * (side == 0 ? -1.0 : 1.0)) / K.Del +
                      c * (1.0 + Ma) * (pres - K.pres_in) / K.Del;
        }
        break;
    case MFC5_BC_CHAR_NR_SUB_OUTFLOW:
        L[0] = L1base; fill(1.0, lam2); L[last] = 0.0;
        if (K.grcbc_out) {
            L[last] = c * (1.0 - Ma) * (pres - K.pres_out) / K.Del;
            if (K.grcbc_vel_out)
                L[last] = L[last] + rho * sq(c) * (1.0 - Ma) * (vel[d] + K.vel_out[d] * (side == 0 ? -1.0 : 1.0)) / K.Del;
        }
        break;
    case MFC5_BC_CHAR_FF_SUB_OUTFLOW: L[0] = L1base; fill(1.0, lam2); L[last] = L[0] + 2.0 * rho * c * lam2 * dvel_ds[d]; break;
    case MFC5_BC_CHAR_CP_SUB_OUTFLOW: L[0] = L1base; fill(1.0, lam2); L[last] = -L[0]; break;
    case MFC5_BC_CHAR_SUP_INFLOW: break;
    case MFC5_BC_CHAR_SUP_OUTFLOW: L[0] = L1base; fill(1.0, lam2); L[last] = lam3 * (dpres_ds + rho * c * dvel_ds[d]); break;
    }
    const double dpres_dt = -5.e-1 * (L[last] + L[0]);
    double darho_dt[NF], dvel_dt[3] = {0, 0, 0}, dadv_dt[NF];
    for (int f = 0; f < NF; f++) darho_dt[f] = -(L[1 + f] - mf[f] * dpres_dt) / (c * c);
    for (int ii = 0; ii < ND; ii++) {
        const int a = didx[ii];
        const double flg = (a == d) ? 1.0 : 0.0;
        dvel_dt[a] = flg * (L[0] - L[last]) / (2.0 * rho * c) + (flg - 1.0) * L[S::IM + ii];
    }
    double vel_dv_dt_sum = 0.0;
    for (int a = 0; a < ND; a++) vel_dv_dt_sum = vel_dv_dt_sum + vel[a] * dvel_dt[a];
    for (int f = 0; f < NF; f++) dadv_dt[f] = -L[S::IE + f];
    double drho_dt = 0, dgamma_dt = 0, dpi_inf_dt = 0, dqv_dt = 0;
    for (int f = 0; f < NF; f++) {
        drho_dt = drho_dt + darho_dt[f];
        dgamma_dt = dgamma_dt + dadv_dt[f] * P.gam[f];
        dpi_inf_dt = dpi_inf_dt + dadv_dt[f] * P.pinf[f];
        dqv_dt = dqv_dt + darho_dt[f] * P.qv[f];
    }
    const double ds0 = K.ds0;
    for (int f = 0; f < NF; f++) fl[0][f] = fl[1][f] + ds0 * darho_dt[f];
    for (int a = 0; a < ND; a++) fl[0][S::IM + a] = fl[1][S::IM + a] + ds0 * (vel[a] * drho_dt + rho * dvel_dt[a]);
    fl[0][S::IE] = fl[1][S::IE] + ds0 * (pres * dgamma_dt + gam * dpres_dt + dpi_inf_dt + dqv_dt + rho * vel_dv_dt_sum +
                                         5.e-1 * drho_dt * vel_K_sum);
    if (hll1) {
        for (int f = 0; f < NF; f++) fl[0][S::IA + f] = 0.0;
        for (int f = 0; f < NF; f++)
            fsl[0][f] = 1.0 / fmax(fabs(vel[d]), MFC5_SGM_EPS) * fsign(1.0, vel[d]) *
                        (fl[1][S::IA + f] + vel[d] * fsl[1][f] + ds0 * dadv_dt[f]);
    } else {
        for (int f = 0; f < NF; f++) fl[0][S::IA + f] = fl[1][S::IA + f] + ds0 * dadv_dt[f];
        for (int f = 0; f < NF; f++) fsl[0][f] = fsl[1][f];
    }
    // store with the sign conventions of s_finalize_cbc; slot jl+1, comp-major, plane-minor
    double* out = P.cbcbuf[d][side];
    const int ncomp = P.ncbc_comp;
#pragma unroll
    for (int jl = -1; jl <= 1; jl++) {
        if (jl > nface - 1) break;
#pragma unroll
        for (int v = 0; v < N; v++) out[((long long)(jl + 1) * ncomp + v) * plane + t] = fl[jl + 1][v] * (v == S::IM + d ? 1.0 : sgn);
        if (hll1) for (int f = 0; f < NF; f++) out[((long long)(jl + 1) * ncomp + N + f) * plane + t] = fsl[jl + 1][f];
        else out[((long long)(jl + 1) * ncomp + N) * plane + t] = fsl[jl + 1][0] * sgn;
    }
}

// the characteristic-BC faces face0 .. of the block in ONE launch (face = 2 * direction + side); Kall[6] in HBM
#ifndef MFC5_CBC_MINB
#define MFC5_CBC_MINB 3   // resident blocks per SM the register allocation aims at: 168 registers + 1 kB of spills; 1 / 3 / 4: x faces 25.7 / 21.5 / 27.6 us, y + z faces 39.5 / 30.0 / 28.4 us (profiles/variants_r02.txt)
#endif
template <int NF, int ND>
__global__ void __launch_bounds__(128, MFC5_CBC_MINB) k_cbc(const __grid_constant__ DevP P, const CbcCoef* __restrict__ Kall, int face0) {
    const int face = blockIdx.y + face0;
    const int d = face >> 1, side = face & 1;
    if (d >= ND || !P.cbc_on[d][side]) return;
    const CbcCoef& K = Kall[face];
    if (d == 0) cbc_point<NF, ND, 0>(P, K, side);
    else if (d == 1) cbc_point<NF, ND, (ND > 1 ? 1 : 0)>(P, K, side);
    else cbc_point<NF, ND, (ND > 2 ? 2 : 0)>(P, K, side);
}

// -------------------------------------------------------------------------------------------------
// shared tail of both sweep kernels: given the fluxes of the two faces of cell `c`, apply the flux
// difference (m_rhs.fpp:1107-1120 / :1183-1196 / :1345-1358), the alpha source (:1440-1488) and,
// when this is the last direction of a time-stepping stage, the fused TVD-RK update
// (m_time_steppers.fpp:486-507).
// -------------------------------------------------------------------------------------------------
// Opt-in (MFC5_FUSE_C2P=1): the stage that just produced the new conservative state of cell c also writes its primitive
// variables into the OTHER primitive buffer, so the next stage starts without a separate cons -> prim pass (N doubles per
// cell less HBM traffic, one launch less per stage).  Same arithmetic as k_cons2prim, so the bits are the same.  Measured
// on C5: z sweep 0.955 -> 1.150 ms against 0.215 ms of k_cons2prim saved, 8.95 -> 8.88 ms per step, for a fifth state-sized
// buffer -- not worth it by default (a reciprocal-based conversion gets 8.80 ms but costs 1e-11 of parity).
// body forces, operator-split after the stage update of one cell (s_apply_bodyforces m_time_steppers.fpp:713-737 with
// s_compute_body_forces_rhs m_body_forces.fpp:294-385): rhoM and the momenta of the UPDATED state enter the source
template <int NF, int ND>
__device__ __forceinline__ void apply_body_forces(const DevP& P, double* qn) {
    using S = Shape<NF, ND>;
    double rhoM = 0.0, rhs_E = 0.0, rhs_m[ND];
#pragma unroll
    for (int f = 0; f < NF; f++) rhoM = rhoM + qn[f];
#pragma unroll
    for (int d = 0; d < ND; d++) {
        rhs_m[d] = rhoM * P.bf_a[d];
        rhs_E = rhs_E + qn[S::IM + d] * P.bf_a[d];
    }
#pragma unroll
    for (int d = 0; d < ND; d++) qn[S::IM + d] = qn[S::IM + d] + P.bf_ldt * rhs_m[d];
    qn[S::IE] = qn[S::IE] + P.bf_ldt * rhs_E;
}

template <int NF, int ND, int FAST>
__device__ __forceinline__ void store_next_prim(const DevP& P, long long c, const double* qn) {
    if (!P.prim_next) return;
    double pr[Shape<NF, ND>::N];
    cons2prim_cell<NF, ND>(P, qn, pr);
#pragma unroll
    for (int v = 0; v < Shape<NF, ND>::N; v++) P.prim_next[v * P.fs + c] = pr[v];
}

template <int NF, int ND, int DIR, int FAST, class ix = long long>
__device__ __forceinline__ void finish_cell(const DevP& P, ix c, double ds, double inv_ds_tab, const double* Fp, const double* Fsp,
                                            const double* Fc, const double* Fsc, const double* Fvp = nullptr, const double* Fvc = nullptr) {
    const double inv_ds = FAST ? inv_ds_tab : 1.0 / ds;  // FAST: 1/ds from the table (host division, correctly rounded like the reference's)
    using S = Shape<NF, ND>;
    const ix fs = (ix)P.fs;
    const bool alpha_iface = (P.rs == 1 && !P.hll_u);
    double r[S::N];
#pragma unroll
    for (int v = 0; v < S::N; v++) {
        if (DIR == 0) r[v] = inv_ds * (Fp[v] - Fc[v]);
        else r[v] = P.rhs[(ix)((ix)v * fs + c)] + inv_ds * (Fp[v] - Fc[v]);
    }
    if (alpha_iface) {
        const double coeff = __ldg(P.prim + (ix)((ix)(S::IM + DIR) * fs + c));
#pragma unroll
        for (int f = 0; f < NF; f++) r[S::IA + f] = r[S::IA + f] + inv_ds * coeff * (Fsp[f] - Fsc[f]);
    } else {
#pragma unroll
        for (int f = 0; f < NF; f++) {
            const double coeff = __ldg(P.prim + (ix)((ix)(S::IA + f) * fs + c));  // q_cons(adv) == q_prim(adv)
            r[S::IA + f] = r[S::IA + f] + inv_ds * coeff * (Fsc[0] - Fsp[0]);
        }
    }
    if (!FAST && P.alt_ss) {  // K div(u) correction, u-interface form (m_rhs.fpp:1489-1503)
        const double K = k_term(P, __ldg(P.prim + (ix)((ix)S::IE * fs + c)), __ldg(P.prim + (ix)((ix)S::IA * fs + c)), __ldg(P.prim + (ix)((ix)(S::IA + NF - 1) * fs + c)));
        r[S::IA] = r[S::IA] + K * inv_ds * (Fsc[0] - Fsp[0]);
        r[S::IA + NF - 1] = r[S::IA + NF - 1] - K * inv_ds * (Fsc[0] - Fsp[0]);
    }
    if (!FAST && Fvp) {   // viscous source fluxes, s_compute_additional_physics_rhs m_rhs.fpp:1762-1770 (after the advection source terms)
#pragma unroll
        for (int a = 0; a <= ND; a++) r[S::IM + a] = r[S::IM + a] + inv_ds * (Fvp[a] - Fvc[a]);
    }
    if (DIR == ND - 1 && P.do_rk) {
        double qn[S::N];
#pragma unroll
        for (int v = 0; v < S::N; v++) {
            const double q1 = P.q1[(ix)((ix)v * fs + c)];
            const double q2 = (P.rk1 != 0.0) ? P.q2[(ix)((ix)v * fs + c)] : 0.0;  // rk1 = 0 in stage 1: 0 * q2 adds nothing
            if (FAST) qn[v] = div_by_const(P.rk0 * q1 + P.rk1 * q2 + P.rk2 * P.dt * r[v], P.rk3, P.irk3);
            else qn[v] = (P.rk0 * q1 + P.rk1 * q2 + P.rk2 * P.dt * r[v]) / P.rk3;
        }
        if (P.bf_on) apply_body_forces<NF, ND>(P, qn);
#pragma unroll
        for (int v = 0; v < S::N; v++) P.qout[(ix)((ix)v * fs + c)] = qn[v];
        store_next_prim<NF, ND, FAST>(P, (long long)c, qn);
    } else {
#pragma unroll
        for (int v = 0; v < S::N; v++) P.rhs[(ix)((ix)v * fs + c)] = r[v];
    }
}

// replacement of a Riemann flux by the characteristic-BC flux on the faces the reference rewrites
template <int NF, int ND, int DIR>
__device__ __forceinline__ void cbc_override(const DevP& P, int face, long long plane_idx, long long plane, double* F, double* Fs) {
    using S = Shape<NF, ND>;
    const int s = P.s[DIR];
    const bool hll1 = (P.rs == 1 && !P.hll_u);
    int slot = -1, side = 0;
    if (face > s) return;
    if (P.cbc_on[DIR][0] && face <= P.nface - 1) { slot = face + 1; side = 0; }
    if (P.cbc_on[DIR][1] && face >= s - P.nface) { slot = (s - 1 - face) + 1; side = 1; }  // local jl = s-1-face
    if (slot < 0) return;
    const double* buf = P.cbcbuf[DIR][side];
    const int ncomp = P.ncbc_comp;
#pragma unroll
    for (int v = 0; v < S::N; v++) F[v] = buf[((long long)slot * ncomp + v) * plane + plane_idx];
    if (hll1) {
#pragma unroll
        for (int f = 0; f < NF; f++) Fs[f] = buf[((long long)slot * ncomp + S::N + f) * plane + plane_idx];
    } else Fs[0] = buf[((long long)slot * ncomp + S::N) * plane + plane_idx];
}

// -------------------------------------------------------------------------------------------------
// x-sweep.  A warp owns cells [i0-1, i0+30] of one x-row: lane l reconstructs cell i0-1+l, receives
// the left-edge state of cell +1 by __shfl_down, solves the Riemann problem of its right face,
// receives the flux of its left face by __shfl_up; lanes 1..30 finish their cell.
// -------------------------------------------------------------------------------------------------
#define MFC5_XW 30  // cells finished per warp
#ifndef MFC5_WENO_NV
#define MFC5_WENO_NV 2  // variables reconstructed per iteration of the rolled WENO loop
#endif

#ifndef MFC5_X_MINB
#define MFC5_X_MINB 4   // 128 registers: 16 warps per SM (measured best of 3, 4, 5 blocks)
#endif
#ifndef MFC5_X_RPB
#define MFC5_X_RPB 4    // rows a warp handles in turn; measured 4: 0.818, 8: 0.823, 16: 0.838, 32: 0.887 ms (the next row's window is in flight while the current one is computed)
#endif
#ifndef MFC5_X_VEC
#define MFC5_X_VEC 1    // 1: segments whose stencil window is contiguous in memory stage it with 16-byte cp.async (20 lanes x double2 per variable
                        // and row instead of 36 8-byte copies); the window then starts one cell earlier, on a 16-byte boundary
#endif
#ifndef MFC5_X_STAB
#define MFC5_X_STAB 1   // 1: the segment's WENO table tile is staged in shared memory once per block
#endif
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// One segment's tile of DirTables::tabX in shared memory: element k of lane l at tile[(k / 2) * 64 + 2 * l + (k & 1)]
// (pairs of consecutive entries are adjacent, [k / 2][lane][2], so that the compiler fetches two of them per 16-byte LDS)
struct RowTile {
    static constexpr bool kAlwaysTable = false;
    static constexpr bool kHasMapTerms = true;
    const double* p;   // tile + 2 * lane
    __device__ __forceinline__ double operator[](int k) const { return p[(k >> 1) * 64 + (k & 1)]; }
};

// Grid: blockIdx.x = segment, blockIdx.z = k, blockIdx.y = group of MFC5_X_RPB * (warps per block) rows j; warp w of the
// block takes rows j0 + w, j0 + w + nw, ... (no integer division anywhere).  All warps of a block work on the same segment,
// so its 48 x 32 table tile is read from HBM / L2 once per block and then from shared memory.  blockDim.x = 128, or 32 for
// thin domains (sweep_grid_x).
template <int NF, int ND, int WO, int WV, int RS, int FAST>
__global__ void __launch_bounds__(128, MFC5_X_MINB) k_sweep_x(const __grid_constant__ DevP P) {
    using S = Shape<NF, ND>;
    constexpr int N = S::N;
    constexpr int XT = 40;
    constexpr int NV = MFC5_WENO_NV;
    constexpr int PNX = (WO == 7) ? 3 : 2;   // half-width of the staged stencil window (the generic kernels reserve the WENO5 one)
    constexpr int FW = 2 * PNX + 1;
    constexpr int NB = (MFC5_X_RPB > 1 && N <= 8) ? 2 : 1;  // window buffers per warp (static shared memory: 48 KB)
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int seg = P.xseg_lo + (int)blockIdx.x;
    if (seg >= P.xseg_hole) seg += P.xseg_hole_n;
    const int k = blockIdx.z;
    const long long rows = (long long)(P.s[1] + 1) * (P.s[2] + 1);
    const int i0 = seg * MFC5_XW;
    const int i = i0 - 1 + lane;              // my cell
    const int ic = min(i, P.s[0] + 1);        // clamp loads (cells beyond m+1 are never used)
    const DirTables& T = P.T[0];
    const int slot = ic + P.b;
    // the warp's 36-cell window of every variable is staged in shared memory with cp.async; lane l reads cells l .. l+4
    __shared__ __align__(16) double sx[4][NB][N][XT];
    __shared__ __align__(16) double2 sedge[4][N][32];  // per lane: (.x right edge of my cell, .y left edge of cell +1): one 16-byte access each way (spill space of the rolled loop)
    __shared__ __align__(16) double stab[(FAST && MFC5_X_STAB) ? MFC5_TAB * 32 : 32];
    double2(*myE)[32] = sedge[w];
    const int hi = P.s[0] + P.b;
    // window cells of this lane (clamped to the ghost region), through the ghost modes of the x boundaries (DevP::gmode)
    bool fa, fb;
    const GhostMap gmap(P, 0);
    const int ca = gmap.src(min(i0 - 1 - PNX + lane, hi), fa), cb = gmap.src(min(i0 - 1 - PNX + 32 + lane, hi), fb);
    // 16-byte staging: the window buffer holds cells i0 - 2 - PNX .. (shifted by SH = 1: i0 = 30 seg and interior cell 0 sits on a
    // 128-byte boundary, so cell i0 - 4 is 16-byte aligned), 40 cells = 20 double2, when all of them are real memory cells of this
    // row (block-uniform: not for segments that touch a folded boundary or the end of the ghost region)
    constexpr int SH = (MFC5_X_VEC && PNX == 2) ? 1 : 0;
    const int wstart = i0 - 1 - PNX - SH;
    const bool vec = SH == 1 && wstart >= (P.gmode[0][0] == MFC5_GM_MEM ? -P.b : 0) &&
                     wstart + 39 <= (P.gmode[0][1] == MFC5_GM_MEM ? hi : P.s[0]);
    const bool any_flip = (P.gmode[0][0] == MFC5_GM_MIRROR) | (P.gmode[0][1] == MFC5_GM_MIRROR);   // kernel-uniform
    using ix = typename SweepIdx<(WO >= 0 && FAST)>::type;  // 32-bit element indices in the case-optimised kernels (k_sweep_march)
    const ix fs = (ix)P.fs;
    auto load_window = [&](int buf, int jj) {
        const ix rb = (ix)P.base + (ix)jj * (ix)P.sy + (ix)k * (ix)P.sz;
        if (vec) {
            if (lane < 20) {
#pragma unroll
                for (int v = 0; v < N; v++) cp_async16(&sx[w][buf][v][2 * lane], P.prim + (ix)((ix)v * fs + rb + (ix)(wstart + 2 * lane)));
            }
        } else {
#pragma unroll
            for (int v = 0; v < N; v++) {
                cp_async8(&sx[w][buf][v][lane + SH], P.prim + (ix)((ix)v * fs + rb + (ix)ca));
                if (lane < 2 * PNX) cp_async8(&sx[w][buf][v][32 + lane + SH], P.prim + (ix)((ix)v * fs + rb + (ix)cb));
            }
        }
        cp_async_commit();
    };
    int j = (int)blockIdx.y * (MFC5_X_RPB * nw) + w;
    if (j <= P.s[1]) load_window(0, j);  // warp-uniform
    if (FAST && MFC5_X_STAB) {           // every thread of the block takes part before any leaves
        const double* src = T.tabX + (long long)seg * (MFC5_TAB * 32);
        for (int q = threadIdx.x; q < MFC5_TAB * 32; q += blockDim.x) stab[q] = __ldg(src + q);
        __syncthreads();
    }
#if MFC5_X_STAB
    const RowTile xrow{stab + 2 * lane};
#else
    const RowSeg xrow{T.tabX + (long long)seg * (MFC5_TAB * 32) + 2 * lane};
#endif
    const double ds_i = __ldg(T.ds + ic + P.b);
#pragma unroll 1
    for (int it = 0; it < MFC5_X_RPB; it++, j += nw) {
        if (j > P.s[1]) break;  // warp-uniform; no block-wide synchronisation below
        const int buf = it & (NB - 1);
        if (NB > 1 && it + 1 < MFC5_X_RPB && j + nw <= P.s[1]) {
            load_window(buf ^ 1, j + nw);  // the next row's window travels while this one is computed
            cp_async_wait_1();
        } else {
            if (NB == 1 && it > 0) load_window(0, j);
            cp_async_wait_all();
        }
        __syncwarp();
        double(*my)[XT] = sx[w][buf];
        if (any_flip) {   // reflective wall: the mirrored cells carry -u_x
            if (fa) my[S::IM][lane + SH] = -my[S::IM][lane + SH];
            if (lane < 2 * PNX && fb) my[S::IM][32 + lane + SH] = -my[S::IM][32 + lane + SH];
            __syncwarp();
        }
        const long long row = j + (long long)(P.s[1] + 1) * k;
        const ix rowbase = (ix)P.base + (ix)j * (ix)P.sy + (ix)k * (ix)P.sz;
        // WENO of NV variables per iteration of a ROLLED loop (code size: the instruction cache, not the
        // register file, is what the fully unrolled version ran out of); the left edge of cell i+1 arrives by
        // warp shuffle; edge states wait in shared memory for the Riemann solve
#pragma unroll 1
        for (int v0 = 0; v0 < N; v0 += NV) {
            double f[NV][FW], l[NV], r[NV];
            int vv[NV];
#pragma unroll
            for (int u = 0; u < NV; u++) {
                vv[u] = min(v0 + u, N - 1);
#pragma unroll
                for (int q = 0; q < FW; q++) f[u][q] = my[vv[u]][lane + q + SH];
            }
            if constexpr (WO == 7) {
#pragma unroll
                for (int u = 0; u < NV; u++) weno7_edges(P, T, slot, f[u], l[u], r[u]);
            } else {
                weno_edges_nv<WO, WV, FAST, NV>(P, T, slot, xrow, f, l, r);
            }
#pragma unroll
            for (int u = 0; u < NV; u++) {
                myE[vv[u]][lane] = make_double2(r[u], __shfl_down_sync(0xffffffffu, l[u], 1));
            }
        }
        if (!FAST && WO < 0 && NF >= 2 && P.int_comp == 1) {  // THINC on my cell: my right edge, and my left edge = lane-1's qR
            double tl[4], tr[4];
            const bool on = thinc_cell(P, my[S::IA][lane + 1 + SH], my[S::IA][lane + 2 + SH], my[S::IA][lane + 3 + SH], my[0][lane + 2 + SH], my[NF - 1][lane + 2 + SH], tl, tr);
            __syncwarp();
            if (on) {
                const int vi[4] = {0, NF - 1, S::IA, S::IA + NF - 1};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    myE[vi[q]][lane].x = tr[q];
                    if (lane > 0) myE[vi[q]][lane - 1].y = tl[q];
                }
            }
            __syncwarp();
        }
        double eR[N], qR[N];
#pragma unroll
        for (int v = 0; v < N; v++) { const double2 e = myE[v][lane]; eR[v] = e.x; qR[v] = e.y; }
        // BC -4 (Riemann extrapolation), m_riemann_state.fpp:149-206 (the case-optimised kernels are not dispatched for it)
        constexpr bool kExtrap = !(FAST && WO >= 0);
        if (kExtrap && P.bc[0][0] == MFC5_BC_RIEMANN_EXTRAP && i == -1) {
#pragma unroll
            for (int v = 0; v < N; v++) eR[v] = qR[v];
        }
        if (kExtrap && P.bc[0][1] == MFC5_BC_RIEMANN_EXTRAP && i == P.s[0]) {
#pragma unroll
            for (int v = 0; v < N; v++) qR[v] = eR[v];
        }
        double F[N], Fs[NF], Fv[FAST ? 1 : ND + 1], Fvp[FAST ? 1 : ND + 1];
#pragma unroll
        for (int f = 0; f < NF; f++) Fs[f] = 0.0;
#pragma unroll
        for (int a = 0; a < (FAST ? 1 : ND + 1); a++) Fv[a] = 0.0;
#ifdef MFC5_DEBUG_SPLIT   // accuracy bisection builds: 1 = reference-order Riemann solve with the fast WENO
        if (FAST && MFC5_DEBUG_SPLIT != 1) riemann_face_fast<NF, ND, 0, RS>(P, eR, qR, F, Fs);
#else
        if (FAST) riemann_face_fast<NF, ND, 0, RS>(P, eR, qR, F, Fs);
#endif
        else if (!FAST && P.viscous) {   // the face between my cell i and cell i + 1 (lanes beyond m: never finished, loads clamped)
            VisFace vo;
            riemann_face<NF, ND, 0, RS>(P, eR, qR, F, Fs, &vo);
            const int idx[3] = {min(i, P.s[0]), j, k};
            viscous_face<NF, ND, 0>(P, (long long)P.base + idx[0] + (long long)j * P.sy + (long long)k * P.sz, idx, vo, Fv);
        }
        else riemann_face<NF, ND, 0, RS>(P, eR, qR, F, Fs);
        if (P.cbc_on[0][0] | P.cbc_on[0][1]) cbc_override<NF, ND, 0>(P, i, row, rows, F, Fs);
        double Fp[N], Fsp[NF];
#pragma unroll
        for (int v = 0; v < N; v++) Fp[v] = __shfl_up_sync(0xffffffffu, F[v], 1);
#pragma unroll
        for (int f = 0; f < NF; f++) Fsp[f] = __shfl_up_sync(0xffffffffu, Fs[f], 1);
        const bool visc = !FAST && P.viscous;   // kernel-uniform
        if (visc) {
#pragma unroll
            for (int a = 0; a < (FAST ? 1 : ND + 1); a++) Fvp[a] = __shfl_up_sync(0xffffffffu, Fv[a], 1);
        }
        if (lane >= 1 && lane <= MFC5_XW && i <= P.s[0]) {
            finish_cell<NF, ND, 0, FAST, ix>(P, rowbase + (ix)i, ds_i, FAST ? xrow[46] : 0.0, Fp, Fsp, F, Fs, visc ? Fvp : nullptr, Fv);
        }
        __syncwarp();  // the window buffer and the edge arrays are rewritten by the next rows
    }
}

// -------------------------------------------------------------------------------------------------
// Marching sweep (all three directions).  Each thread owns one pencil of cells along DIR and marches a
// chunk [c0, c1] of it; the 128 threads of a block are 128 neighbouring pencils:
//     DIR = 1, 2 : lanes run along x (coalesced loads / stores),
//     DIR = 0    : lanes run along y; a lane's 8-byte accesses to consecutive x fill 32-byte sectors
//                  over 4 successive iterations out of L1 / L2.
// Data staging: the (2*PN+1)-plane WENO stencil window of the primitive variables lives in a
// per-thread ring in SHARED MEMORY (layout [slot][var][lane], conflict free, no block barrier needed:
// a lane only ever touches its own column).  The plane entering the window is requested one iteration
// ahead with cp.async (LDGSTS, global -> shared without register staging), so the HBM latency is
// covered by a whole iteration of arithmetic; the inputs of the finishing step are pulled into L2.
// Right-edge states cross iterations through shared memory as well; face states and face fluxes never
// touch HBM.  Iteration cc reconstructs cell cc, solves face cc-1 with the stored right edge of cell
// cc-1, finishes cell cc-1 with the carried flux of face cc-2 (flux difference, alpha source, fused RK).
// The WENO tables are indexed by the marching position, i.e. uniform over the block.
// -------------------------------------------------------------------------------------------------
#ifndef MFC5_PF_MASK
#define MFC5_PF_MASK 7   // L2 prefetches of the marching kernel: 1 = next primitive plane, 2 = rhs, 4 = q1 / q2 of the cell to finish
#endif
#ifndef MFC5_MARCH_THREADS
#define MFC5_MARCH_THREADS 192
#endif
#ifndef MFC5_MARCH_MINB
#define MFC5_MARCH_MINB 2   // resident blocks per SM the marching kernel is built for (N <= 8)
#endif

template <int WO>
struct MarchCfg {
    static constexpr int PN = WO < 0 ? 2 : (WO - 1) / 2;  // generic kernels reserve the WENO5 window
    static constexpr int R = 2 * PN + 1;   // stencil window
    static constexpr int NSLOT = R;        // ring slots: the window; the entering plane replaces the leaving one
};

// dynamic shared memory of a block marching a chunk of `chunk` cells: stencil ring + right-edge ping-pong
// + left edges + the chunk's rows of the WENO table
template <int NF, int ND, int WO>
inline size_t march_smem_bytes(int chunk) {
    return ((size_t)(MarchCfg<WO>::NSLOT + 3) * (2 * NF + ND + 1) * MFC5_MARCH_THREADS + (size_t)(chunk + 2) * MFC5_TAB) * sizeof(double);
}

template <int NF, int ND, int DIR, int WO, int WV, int RS, int FAST>
__global__ void __launch_bounds__(MFC5_MARCH_THREADS, (2 * NF + ND + 1 <= 8) ? MFC5_MARCH_MINB : 1) k_sweep_march(const __grid_constant__ DevP P) {
    using S = Shape<NF, ND>;
    constexpr int N = S::N;
    constexpr int PN = MarchCfg<WO>::PN, R = MarchCfg<WO>::R, NSLOT = MarchCfg<WO>::NSLOT;
    constexpr int NT = MFC5_MARCH_THREADS;
    extern __shared__ double smem[];
    double* ring = smem;                  // [NSLOT][N][NT]
    double* sER = smem + NSLOT * N * NT;  // [2][N][NT]  right-edge states, ping-pong over iterations
    double* sEL = sER + 2 * N * NT;       // [N][NT]     left-edge states of the cell reconstructed in this iteration
    double* srow = sEL + N * NT;          // [chunk+2][MFC5_TAB]  WENO table rows of cells c0-1 .. c1+1
    const int tid = threadIdx.x;
    if (FAST) {  // cooperative, coalesced copy; every thread of the block takes part before any returns
        const int s_ = P.s[DIR];
        const int c0_ = blockIdx.y * P.chunk, c1_ = min(c0_ + P.chunk - 1, s_);
        const double* src = P.T[DIR].tab + (long long)(c0_ - 1 + P.b) * MFC5_TAB;
        const int nrow = (c1_ - c0_ + 3) * MFC5_TAB;
        for (int k = tid; k < nrow; k += NT) srow[k] = __ldg(src + k);
        __syncthreads();
    }

    const int n_a = (DIR == 0) ? P.s[1] + 1 : P.s[0] + 1;
    const int n_b = (DIR == 2) ? P.s[1] + 1 : P.s[2] + 1;
    const long long plane = (long long)n_a * n_b;
    const long long t = blockIdx.x * (long long)NT + tid;
    if (t >= plane) return;
    const int a = (int)(t % n_a), bb = (int)(t / n_a);
    // element indices into the field arrays: 32-bit (one IMAD.WIDE per access, the per-variable index shared by all arrays)
    // in the case-optimised kernels, whose launcher checks that N * field size fits; 64-bit in the generic ones.  Plane
    // numbers below zero wrap around in unsigned arithmetic and come out right because the final index is non-negative.
    using ix = typename SweepIdx<(WO >= 0 && FAST)>::type;
    const ix fs = (ix)P.fs;
    const ix sa = (DIR == 0) ? (ix)P.sy : (ix)1;
    const ix sb = (DIR == 2) ? (ix)P.sy : (ix)P.sz;
    const ix sd = (DIR == 0) ? (ix)1 : (DIR == 1 ? (ix)P.sy : (ix)P.sz);
    const ix colbase = (ix)P.base + (ix)a * sa + (ix)bb * sb;
    const int s = P.s[DIR];
    const int c0 = blockIdx.y * P.chunk;
    const int c1 = min(c0 + P.chunk - 1, s);
    const DirTables& T = P.T[DIR];
    const int lo = -P.b, hi = s + P.b;  // planes outside the ghost region are never used; clamp the loads

    // prologue: the window of cell c0-1, planes c0-1-PN .. c0-1+PN -> slots 0 .. R-1
#pragma unroll
    for (int k = 0; k < R; k++) {
        const int pl = min(max(c0 - 1 - PN + k, lo), hi);
#pragma unroll
        for (int v = 0; v < N; v++) cp_async8(ring + (k * N + v) * NT + tid, P.prim + (ix)((ix)v * fs + colbase + (ix)pl * sd));
    }
    cp_async_commit();
    int tail = 0;  // slot of the oldest plane of the window (cc - PN)

    double Fp[N], Fsp[NF];  // flux of the previous face
#pragma unroll
    for (int v = 0; v < N; v++) Fp[v] = 0.0;
#pragma unroll
    for (int f = 0; f < NF; f++) Fsp[f] = 0.0;
    constexpr int NVIS = FAST ? 1 : ND + 1;
    double Fvp[NVIS];       // viscous source flux of the previous face (reference-order kernels, viscous = T)
#pragma unroll
    for (int q = 0; q < NVIS; q++) Fvp[q] = 0.0;
    const bool visc = !FAST && P.viscous;

    constexpr int NV = MFC5_WENO_NV;
    for (int cc = c0 - 1; cc <= c1 + 1; cc++) {
        cp_async_wait_all();  // the window of cell cc is complete
        if ((MFC5_PF_MASK & 1) && cc < c1) {        // the plane requested after the NEXT reconstruction: pull it into L2 now
            const int pl = min(max(cc + 2 + PN, lo), hi);
#pragma unroll
            for (int v = 0; v < N; v++) prefetch_l2(P.prim + (ix)((ix)v * fs + colbase + (ix)pl * sd));
        }
        if (cc >= c0 + 1) {   // inputs of the finishing step of cell cc-1: pull them into L2 now
            const ix c = colbase + (ix)(cc - 1) * sd;
            if ((MFC5_PF_MASK & 2) && DIR != 0) {
#pragma unroll
                for (int v = 0; v < N; v++) prefetch_l2(P.rhs + (ix)((ix)v * fs + c));
            }
            if ((MFC5_PF_MASK & 4) && (DIR == ND - 1) && P.do_rk) {
#pragma unroll
                for (int v = 0; v < N; v++) prefetch_l2(P.q1 + (ix)((ix)v * fs + c));
                if (P.rk1 != 0.0) {  // stage 1 has rk1 = 0: q2 is not read at all
#pragma unroll
                    for (int v = 0; v < N; v++) prefetch_l2(P.q2 + (ix)((ix)v * fs + c));
                }
            }
        }
        // tap k (plane cc-PN+k) sits in slot (tail+k) mod NSLOT
        int off[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            int sl = tail + k;
            if (sl >= NSLOT) sl -= NSLOT;
            off[k] = sl * N * NT + tid;
        }
        double* sERc = sER + ((cc & 1) ? N * NT : 0) + tid;
        const double* sERp = sER + ((cc & 1) ? 0 : N * NT) + tid;
        // WENO of NV variables per iteration of a ROLLED loop (code size; the table row is read once per NV
        // variables); edge states go to this thread's column of shared memory
        {
            const RowAoS mrow{srow + (cc - (c0 - 1)) * MFC5_TAB};
#pragma unroll 1
            for (int v0 = 0; v0 < N; v0 += NV) {
                constexpr int FW = (PN == 3) ? 7 : 5, FC = FW / 2;   // f[u][FC + k] = v(cc + k)
                double f[NV][FW], l[NV], r[NV];
                int vv[NV];
#pragma unroll
                for (int u = 0; u < NV; u++) {
                    vv[u] = min(v0 + u, N - 1);
#pragma unroll
                    for (int k = 0; k < FW; k++) f[u][k] = 0.0;
#pragma unroll
                    for (int k = 0; k < R; k++) f[u][FC - PN + k] = ring[off[k] + vv[u] * NT];
                }
                if constexpr (WO == 7) {
#pragma unroll
                    for (int u = 0; u < NV; u++) weno7_edges(P, T, cc + P.b, f[u], l[u], r[u]);
                } else {
                    weno_edges_nv<WO, WV, FAST, NV>(P, T, cc + P.b, mrow, f, l, r);
                }
#pragma unroll
                for (int u = 0; u < NV; u++) { sEL[vv[u] * NT + tid] = l[u]; sERc[vv[u] * NT] = r[u]; }
            }
            if (!FAST && WO < 0 && NF >= 2 && PN >= 1 && P.int_comp == 1) {  // THINC on cell cc (a thread only touches its own column)
                double tl[4], tr[4];
                const int om = off[PN >= 1 ? PN - 1 : 0], oc = off[PN], op = off[PN >= 1 ? PN + 1 : 0];
                if (thinc_cell(P, ring[om + S::IA * NT], ring[oc + S::IA * NT], ring[op + S::IA * NT], ring[oc], ring[oc + (NF - 1) * NT], tl, tr)) {
                    const int vi[4] = {0, NF - 1, S::IA, S::IA + NF - 1};
#pragma unroll
                    for (int q = 0; q < 4; q++) { sEL[vi[q] * NT + tid] = tl[q]; sERc[vi[q] * NT] = tr[q]; }
                }
            }
        }
        if (cc <= c1) {       // plane cc+1+PN enters the window in the slot of plane cc-PN; lands during the Riemann solve
            const int pl = min(max(cc + 1 + PN, lo), hi);
#pragma unroll
            for (int v = 0; v < N; v++) cp_async8(ring + (tail * N + v) * NT + tid, P.prim + (ix)((ix)v * fs + colbase + (ix)pl * sd));
            cp_async_commit();
        }
        if (cc >= c0) {
            const int face = cc - 1;
            const ix c = colbase + (ix)(cc - 1) * sd;  // the cell finished in this iteration
            const bool fin = cc >= c0 + 1;
            // inputs of the finishing step, requested before the Riemann solve so that their latency is covered
            double rin[N], q1in[N], q2in[N], acoef[NF];
            const bool rk = (DIR == ND - 1) && P.do_rk;
            const bool alpha_iface = (P.rs == 1 && !P.hll_u);
            if (fin) {
#pragma unroll
                for (int v = 0; v < N; v++) rin[v] = (DIR == 0) ? 0.0 : P.rhs[(ix)((ix)v * fs + c)];
                if (rk) {   // kernel-uniform branches, outside the variable loop
#pragma unroll
                    for (int v = 0; v < N; v++) q1in[v] = __ldg(P.q1 + (ix)((ix)v * fs + c));
                    if (P.rk1 != 0.0) {
#pragma unroll
                        for (int v = 0; v < N; v++) q2in[v] = __ldg(P.q2 + (ix)((ix)v * fs + c));
                    } else {
#pragma unroll
                        for (int v = 0; v < N; v++) q2in[v] = 0.0;
                    }
                }
#pragma unroll
                for (int f = 0; f < NF; f++) acoef[f] = __ldg(P.prim + (ix)((ix)(alpha_iface ? (S::IM + DIR) : (S::IA + f)) * fs + c));
            }
            double qL[N], eL[N], F[N], Fs[NF], Fv[NVIS];
#pragma unroll
            for (int v = 0; v < N; v++) { qL[v] = sERp[v * NT]; eL[v] = sEL[v * NT + tid]; }
#pragma unroll
            for (int f = 0; f < NF; f++) Fs[f] = 0.0;
#pragma unroll
            for (int q = 0; q < NVIS; q++) Fv[q] = 0.0;
            if (P.bc[DIR][0] == MFC5_BC_RIEMANN_EXTRAP && face == -1) {  // m_riemann_state.fpp:149-206
#pragma unroll
                for (int v = 0; v < N; v++) qL[v] = eL[v];
            }
            if (P.bc[DIR][1] == MFC5_BC_RIEMANN_EXTRAP && face == s) {
#pragma unroll
                for (int v = 0; v < N; v++) eL[v] = qL[v];
            }
#ifdef MFC5_DEBUG_SPLIT
            if (FAST && MFC5_DEBUG_SPLIT != 1) riemann_face_fast<NF, ND, DIR, RS>(P, qL, eL, F, Fs);
#else
            if (FAST) riemann_face_fast<NF, ND, DIR, RS>(P, qL, eL, F, Fs);
#endif
            else if (visc) {   // face cc-1: between cells cc-1 and cc of my pencil
                VisFace vo;
                riemann_face<NF, ND, DIR, RS>(P, qL, eL, F, Fs, &vo);
                int idx[3];
                idx[DIR] = cc - 1;
                idx[DIR == 0 ? 1 : 0] = a;
                idx[DIR == 2 ? 1 : 2] = bb;
                viscous_face<NF, ND, DIR>(P, (long long)c, idx, vo, Fv);
            }
            else riemann_face<NF, ND, DIR, RS>(P, qL, eL, F, Fs);
            if (P.cbc_on[DIR][0] | P.cbc_on[DIR][1]) cbc_override<NF, ND, DIR>(P, face, t, plane, F, Fs);
            if (fin) {
                const double inv_ds = FAST ? srow[(cc - 1 - (c0 - 1)) * MFC5_TAB + 46] : 1.0 / __ldg(T.ds + (cc - 1) + P.b);
                double r[N];
#pragma unroll
                for (int v = 0; v < N; v++) {
                    if (DIR == 0) r[v] = inv_ds * (Fp[v] - F[v]);
                    else r[v] = rin[v] + inv_ds * (Fp[v] - F[v]);
                }
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    if (alpha_iface) r[S::IA + f] = r[S::IA + f] + inv_ds * acoef[f] * (Fsp[f] - Fs[f]);
                    else r[S::IA + f] = r[S::IA + f] + inv_ds * acoef[f] * (Fs[0] - Fsp[0]);
                }
                if (!FAST && P.alt_ss) {  // K div(u) correction, u-interface form (m_rhs.fpp:1568-1582,1714-1728)
                    const double K = k_term(P, __ldg(P.prim + (ix)((ix)S::IE * fs + c)), __ldg(P.prim + (ix)((ix)S::IA * fs + c)), __ldg(P.prim + (ix)((ix)(S::IA + NF - 1) * fs + c)));
                    r[S::IA] = r[S::IA] + K * inv_ds * (Fs[0] - Fsp[0]);
                    r[S::IA + NF - 1] = r[S::IA + NF - 1] - K * inv_ds * (Fs[0] - Fsp[0]);
                }
                if (visc) {   // s_compute_additional_physics_rhs, m_rhs.fpp:1762-1770,1847-1855,1939-1947
#pragma unroll
                    for (int q = 0; q < NVIS; q++) r[S::IM + q] = r[S::IM + q] + inv_ds * (Fvp[q] - Fv[q]);
                }
                if (rk) {
#pragma unroll
                    for (int v = 0; v < N; v++) {
                        if (FAST) r[v] = div_by_const(P.rk0 * q1in[v] + P.rk1 * q2in[v] + P.rk2 * P.dt * r[v], P.rk3, P.irk3);
                        else r[v] = (P.rk0 * q1in[v] + P.rk1 * q2in[v] + P.rk2 * P.dt * r[v]) / P.rk3;
                    }
                    if (P.bf_on) apply_body_forces<NF, ND>(P, r);
#pragma unroll
                    for (int v = 0; v < N; v++) P.qout[(ix)((ix)v * fs + c)] = r[v];
                    if constexpr (WO < 0) store_next_prim<NF, ND, FAST>(P, (long long)c, r);  // opt-in MFC5_FUSE_C2P: generic kernels only
                } else {
#pragma unroll
                    for (int v = 0; v < N; v++) P.rhs[(ix)((ix)v * fs + c)] = r[v];
                }
            }
#pragma unroll
            for (int v = 0; v < N; v++) Fp[v] = F[v];
#pragma unroll
            for (int f = 0; f < NF; f++) Fsp[f] = Fs[f];
#pragma unroll
            for (int q = 0; q < NVIS; q++) Fvp[q] = Fv[q];
        }
        tail = tail + 1;
        if (tail >= NSLOT) tail = 0;
    }
}

// -------------------------------------------------------------------------------------------------
// Marching sweep, second generation (case-optimised configurations only: division-light arithmetic, compile-time scheme, 32-bit
// field indices, no Riemann-extrapolation boundary).  Same algorithm and data flow as k_sweep_march; what changed is
// everything that is not FP64 arithmetic, because an FP64 warp instruction holds the issue port for two cycles and every other
// instruction for one (time ~ 2 F + O per warp; profiles/sass_budget_r02.md):
//   * shared memory per thread: the window ring has ONE extra slot that carries the right-edge states of the previous cell --
//     the right edges of cell cc are written over the plane that leaves the window (cc - PN), variable pair by variable pair
//     as the rolled WENO loop retires it -- instead of a two-slot ping-pong: (R + 2) slots instead of (R + 3);
//   * L2 prefetches are issued per 128-byte line, not per lane: lane l of a warp covers variable l & 7 at the column of one of
//     four sample lanes, so one instruction per array replaces eight;
//   * the inputs of the finishing step (rhs, q1, q2) are loaded after the Riemann solve (they are in L2 by then), which frees
//     ~48 registers across it;
//   * ghost planes of extrapolation / reflective / periodic boundaries are read straight from the interior plane the boundary
//     condition copies (DevP::gmode), so no ghost-fill kernel runs for them;
//   * the RK-fused last sweep converts the new state to primitive variables in registers and stores them over the rhs entry it
//     has just consumed (DevP::prim_next == rhs): the next stage starts without a cons -> prim pass and without a fifth buffer.
// -------------------------------------------------------------------------------------------------
#ifndef MFC5_M2_THREADS
#define MFC5_M2_THREADS 192
#endif
#ifndef MFC5_M2_MINB
#define MFC5_M2_MINB 2
#endif
#ifndef MFC5_M2_TABS
#define MFC5_M2_TABS 1      // 1: the chunk's table rows are staged in shared memory; 0: read through L1 (uniform-address loads)
#endif
#ifndef MFC5_M2_LATE
#define MFC5_M2_LATE 0      // 1: finishing inputs are loaded after the Riemann solve (fewer live registers, exposed L2 latency; measured: z 0.937 vs 0.913 ms)
#endif
#ifndef MFC5_M2_NV
#define MFC5_M2_NV MFC5_WENO_NV   // variables per iteration of the rolled WENO loop
#endif
#ifndef MFC5_M2_DIET
#define MFC5_M2_DIET 0      // 1: fewer address / index instructions per iteration (running field pointers stepped by the field size, the
                            // ghost map of the plane that enters next carried over from its prefetch); A/B in profiles/variants_r02.txt
#endif
#ifndef MFC5_M2_FPS
#define MFC5_M2_FPS 0       // 1: the flux of the previous face crosses iterations through shared memory instead of 16 registers + their copies
#endif

template <int NF, int ND, int WO>
inline size_t march2_smem_bytes(int chunk) {
    constexpr int R = 2 * ((WO - 1) / 2) + 1;
    return ((size_t)(R + 2 + MFC5_M2_FPS) * (2 * NF + ND + 1) * MFC5_M2_THREADS + (MFC5_M2_TABS ? (size_t)(chunk + 2) * MFC5_TAB : 0)) * sizeof(double);
}

// cons -> prim of one cell with reciprocals + residual-corrected quotients (same formulas and order as cons2prim_cell)
template <int NF, int ND>
__device__ __forceinline__ void cons2prim_cell_fast(const DevP& P, const double* q, double* pr) {
    using S = Shape<NF, ND>;
    double arho[NF], alf[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) { arho[f] = q[f]; alf[f] = q[S::IA + f]; }
    if (P.mpp) {
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < NF; i++) {
                arho[i] = sel_max(arho[i], 0.0);
                alf[i] = sel_min(sel_max(alf[i], 0.0), 1.0);
                sum = sum + alf[i];
            }
            const double den = sel_max(sum, MFC5_SGM_EPS), r = rcp_nr(den);
#pragma unroll
            for (int i = 0; i < NF; i++) alf[i] = div_r(alf[i], den, r);
        }
    }
    double rho = 0, gam = 0, pinf = 0, qv = 0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        rho = rho + arho[i];
        gam = fma(alf[i], P.gam[i], gam);
        pinf = fma(alf[i], P.pinf[i], pinf);
        qv = fma(arho[i], P.qv[i], qv);
    }
    rho = sel_max(rho, MFC5_SGM_EPS);
    const double irho = rcp_nr(rho);
    double dyn = 0.0;
#pragma unroll
    for (int a = 0; a < ND; a++) {
        const double u = div_r(q[S::IM + a], rho, irho);
        dyn = fma(5.e-1 * q[S::IM + a], u, dyn);
        pr[S::IM + a] = u;
    }
    pr[S::IE] = div_r(q[S::IE] - dyn - pinf - qv, gam, rcp_nr(gam));
#pragma unroll
    for (int f = 0; f < NF; f++) {
        pr[f] = q[f];
        pr[S::IA + f] = q[S::IA + f];
    }
}

template <int NF, int ND, int DIR, int WO, int WV, int RS>
__global__ void __launch_bounds__(MFC5_M2_THREADS, MFC5_M2_MINB) k_march2(const __grid_constant__ DevP P) {
    using S = Shape<NF, ND>;
    constexpr int N = S::N;
    constexpr int PN = (WO - 1) / 2, R = 2 * PN + 1, NSLOT = R + 1;
    constexpr int NT = MFC5_M2_THREADS;
    constexpr int NV = MFC5_M2_NV;
    static_assert(N == 8, "the line-wise prefetch pattern assumes 8 variables");
    extern __shared__ double smem[];
    double* ring = smem;                   // [NSLOT][N][NT]: R window planes + the right-edge states of the previous cell
    double* sEL = smem + NSLOT * N * NT;   // [N][NT] left-edge states of the cell reconstructed in this iteration
    double* sFp = sEL + N * NT;            // [N][NT] (MFC5_M2_FPS) flux of the previous face
    double* srow = sFp + MFC5_M2_FPS * N * NT;   // [chunk + 2][MFC5_TAB] table rows of cells c0-1 .. c1+1
    const int tid = threadIdx.x;
    const int s = P.s[DIR];
    const int c0 = blockIdx.y * P.chunk;
    const int c1 = min(c0 + P.chunk - 1, s);
    const DirTables& T = P.T[DIR];
    if (MFC5_M2_TABS) {  // cooperative, coalesced copy; every thread of the block takes part before any returns
        const double* src = T.tab + (long long)(c0 - 1 + P.b) * MFC5_TAB;
        const int nrow = (c1 - c0 + 3) * MFC5_TAB;
        for (int k = tid; k < nrow; k += NT) srow[k] = __ldg(src + k);
        __syncthreads();
    }
    const int n_a = (DIR == 0) ? P.s[1] + 1 : P.s[0] + 1;
    const int n_b = (DIR == 2) ? P.s[1] + 1 : P.s[2] + 1;
    const long long plane = (long long)n_a * n_b;
    const long long t = blockIdx.x * (long long)NT + tid;
    const long long tc = min(t, plane - 1);  // out-of-range threads still take part in the warp shuffles below
    const int a = (int)(tc % n_a), bb = (int)(tc / n_a);
    using ix = unsigned;
    const ix fs = (ix)P.fs;
    const ix sa = (DIR == 0) ? (ix)P.sy : (ix)1;
    const ix sb = (DIR == 2) ? (ix)P.sy : (ix)P.sz;
    const ix sd = (DIR == 0) ? (ix)1 : (DIR == 1 ? (ix)P.sy : (ix)P.sz);
    const ix colbase = (ix)P.base + (ix)a * sa + (ix)bb * sb;
    // line-wise L2 prefetch: lane l covers variable l & 7 in the column of sample lane {0, 10, 21, 31}[l >> 3]
    const int lane = tid & 31;
    const int samp = (lane >> 3) == 0 ? 0 : (lane >> 3) == 1 ? 10 : (lane >> 3) == 2 ? 21 : 31;
    const ix pfbase = (ix)(lane & 7) * fs + (ix)__shfl_sync(0xffffffffu, colbase, samp);
    if (t >= plane) return;
    const int lo = -P.b, hi = s + P.b;
    const bool rk = (DIR == ND - 1) && P.do_rk;
    const bool alpha_iface = (RS == 1 && !P.hll_u);
    const GhostMap gmap(P, DIR);

    // plane pl of the primitive variables -> ring slot k (this thread's column)
    auto load_plane_src = [&](int k, int ps, bool flip) {
        const ix c = colbase + (ix)ps * sd;
#if MFC5_M2_DIET
        const double* pp = P.prim + c;
        double* rp = ring + k * N * NT + tid;
#pragma unroll
        for (int v = 0; v < N; v++) {
            if (v == S::IM + DIR && flip) *rp = -*pp;
            else cp_async8(rp, pp);
            pp += fs; rp += NT;
        }
#else
#pragma unroll
        for (int v = 0; v < N; v++) {
            if (v == S::IM + DIR && flip) ring[(k * N + v) * NT + tid] = -P.prim[(ix)((ix)v * fs + c)];   // reflective wall: -u_n of the mirror cell
            else cp_async8(ring + (k * N + v) * NT + tid, P.prim + (ix)((ix)v * fs + c));
        }
#endif
    };
    auto load_plane = [&](int k, int pl) {
        bool flip;
        const int ps = gmap.src(min(max(pl, lo), hi), flip);
        load_plane_src(k, ps, flip);
    };
    auto pf_plane = [&](const double* arr, int pl) { prefetch_l2(arr + (ix)(pfbase + (ix)pl * sd)); };

    // prologue: the window of cell c0-1, planes c0-1-PN .. c0-1+PN -> slots 0 .. R-1, and already the plane that enters next
    // (c0+PN) in slot R, which holds "the previous right edge" from the second iteration on
#pragma unroll
    for (int k = 0; k < NSLOT; k++) load_plane(k, c0 - 1 - PN + k);
    cp_async_commit();
    int tail = 0;  // slot of the oldest plane of the window (cc - PN)

    double Fp[MFC5_M2_FPS ? 1 : N], Fsp[NF];  // flux of the previous face
#pragma unroll
    for (int v = 0; v < (MFC5_M2_FPS ? 1 : N); v++) Fp[v] = 0.0;
#pragma unroll
    for (int f = 0; f < NF; f++) Fsp[f] = 0.0;

#if MFC5_M2_DIET
    bool pf_flip, nxt_flip = false;
    int pf_ps = gmap.src(min(max(c0 + PN, lo), hi), pf_flip), nxt_ps = 0;   // plane cc + 1 + PN of the first iteration (cc = c0 - 1)
#endif
    for (int cc = c0 - 1; cc <= c1 + 1; cc++) {
        cp_async_wait_all();  // the window of cell cc is complete
        {   // pull into L2 what the NEXT iteration will ask for
            bool flip;
#if MFC5_M2_DIET
            nxt_ps = pf_ps; nxt_flip = pf_flip;   // the plane this iteration loads (cc + 1 + PN) was mapped by the previous one's prefetch
            pf_ps = gmap.src(min(max(cc + 2 + PN, lo), hi), pf_flip);
            if (cc < c1) pf_plane(P.prim, pf_ps);
            (void)flip;
#else
            if (cc < c1) pf_plane(P.prim, gmap.src(min(max(cc + 2 + PN, lo), hi), flip));
#endif
            if (cc >= c0 && cc <= c1) {   // the cell finished in the next iteration is cc
                if (DIR != 0) pf_plane(P.rhs, cc);
                if (rk) {
                    pf_plane(P.q1, cc);
                    if (P.rk1 != 0.0) pf_plane(P.q2, cc);
                }
            }
        }
        // tap k (plane cc-PN+k) sits in slot (tail+k) mod NSLOT; the previous right edge in slot tail-1
        int off[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            int sl = tail + k;
            if (sl >= NSLOT) sl -= NSLOT;
            off[k] = sl * N * NT + tid;
        }
        const int slR = (tail == 0 ? NSLOT - 1 : tail - 1);
        {
            const RowAoST mrow{MFC5_M2_TABS ? srow + (cc - (c0 - 1)) * MFC5_TAB : T.tab + (long long)(cc + P.b) * MFC5_TAB};
#pragma unroll 1
            for (int v0 = 0; v0 < N; v0 += NV) {
                double f[NV][5], l[NV], r[NV];
                int vv[NV];
#pragma unroll
                for (int u = 0; u < NV; u++) {
                    vv[u] = min(v0 + u, N - 1);
#pragma unroll
                    for (int k = 0; k < 5; k++) f[u][k] = 0.0;
                    f[u][2] = ring[off[PN] + vv[u] * NT];
                    if (PN >= 1) { f[u][1] = ring[off[PN - 1] + vv[u] * NT]; f[u][3] = ring[off[PN + (PN >= 1 ? 1 : 0)] + vv[u] * NT]; }
                    if (PN >= 2) { f[u][0] = ring[off[0] + vv[u] * NT]; f[u][4] = ring[off[R - 1] + vv[u] * NT]; }
                }
                weno_edges_nv<WO, WV, 1, NV>(P, T, cc + P.b, mrow, f, l, r);
                // the right edge of cell cc replaces plane cc-PN of the same variable, which this iteration has just read for the last time
#pragma unroll
                for (int u = 0; u < NV; u++) { sEL[vv[u] * NT + tid] = l[u]; ring[off[0] + vv[u] * NT] = r[u]; }
            }
        }
        if (cc >= c0) {
            const int face = cc - 1;
            const ix c = colbase + (ix)(cc - 1) * sd;  // the cell finished in this iteration
            const bool fin = cc >= c0 + 1;
            double qL[N], eL[N], F[N], Fs[NF];
#pragma unroll
            for (int v = 0; v < N; v++) { qL[v] = ring[(slR * N + v) * NT + tid]; eL[v] = sEL[v * NT + tid]; }
            if (cc <= c1) {       // plane cc+1+PN enters the window in the slot the previous right edge has just left
#if MFC5_M2_DIET
                load_plane_src(slR, nxt_ps, nxt_flip);
#else
                load_plane(slR, cc + 1 + PN);
#endif
                cp_async_commit();
            }
#if MFC5_M2_LATE != 1
            double rin[N], q1in[N], q2in[N], acoef[NF];
            if (fin) {
#if MFC5_M2_DIET
                {
                    const double* pr = P.rhs + c;
#pragma unroll
                    for (int v = 0; v < N; v++) { rin[v] = (DIR == 0) ? 0.0 : *pr; pr += fs; }
                }
                if (rk && MFC5_M2_LATE == 0) {
                    const double* p1 = P.q1 + c;
#pragma unroll
                    for (int v = 0; v < N; v++) { q1in[v] = __ldg(p1); p1 += fs; }
                    const double* p2 = P.q2 + c;
#pragma unroll
                    for (int v = 0; v < N; v++) { q2in[v] = (P.rk1 != 0.0) ? __ldg(p2) : 0.0; p2 += fs; }
                }
#else
#pragma unroll
                for (int v = 0; v < N; v++) rin[v] = (DIR == 0) ? 0.0 : P.rhs[(ix)((ix)v * fs + c)];
                if (rk && MFC5_M2_LATE == 0) {
#pragma unroll
                    for (int v = 0; v < N; v++) q1in[v] = __ldg(P.q1 + (ix)((ix)v * fs + c));
#pragma unroll
                    for (int v = 0; v < N; v++) q2in[v] = (P.rk1 != 0.0) ? __ldg(P.q2 + (ix)((ix)v * fs + c)) : 0.0;
                }
#endif
#pragma unroll
                for (int f = 0; f < NF; f++) acoef[f] = __ldg(P.prim + (ix)((ix)(alpha_iface ? (S::IM + DIR) : (S::IA + f)) * fs + c));
            }
#endif
#pragma unroll
            for (int f = 0; f < NF; f++) Fs[f] = 0.0;
            riemann_face_fast<NF, ND, DIR, RS>(P, qL, eL, F, Fs);
            if (P.cbc_on[DIR][0] | P.cbc_on[DIR][1]) cbc_override<NF, ND, DIR>(P, face, t, plane, F, Fs);
            if (fin) {
                const double inv_ds = MFC5_M2_TABS ? srow[(cc - 1 - (c0 - 1)) * MFC5_TAB + 46] : __ldg(T.tab + (long long)(cc - 1 + P.b) * MFC5_TAB + 46);
#if MFC5_M2_LATE == 2   // rhs and the alpha coefficients early, q1 / q2 after the Riemann solve
                if (rk) {
#pragma unroll
                    for (int v = 0; v < N; v++) q1in[v] = __ldg(P.q1 + (ix)((ix)v * fs + c));
#pragma unroll
                    for (int v = 0; v < N; v++) q2in[v] = (P.rk1 != 0.0) ? __ldg(P.q2 + (ix)((ix)v * fs + c)) : 0.0;
                }
#endif
#if MFC5_M2_LATE == 1
                double rin[N], q1in[N], q2in[N], acoef[NF];
#pragma unroll
                for (int v = 0; v < N; v++) rin[v] = (DIR == 0) ? 0.0 : P.rhs[(ix)((ix)v * fs + c)];
                if (rk) {
#pragma unroll
                    for (int v = 0; v < N; v++) q1in[v] = __ldg(P.q1 + (ix)((ix)v * fs + c));
#pragma unroll
                    for (int v = 0; v < N; v++) q2in[v] = (P.rk1 != 0.0) ? __ldg(P.q2 + (ix)((ix)v * fs + c)) : 0.0;
                }
#pragma unroll
                for (int f = 0; f < NF; f++) acoef[f] = __ldg(P.prim + (ix)((ix)(alpha_iface ? (S::IM + DIR) : (S::IA + f)) * fs + c));
#endif
                double r[N];
#pragma unroll
                for (int v = 0; v < N; v++) {
                    const double fp = MFC5_M2_FPS ? sFp[v * NT + tid] : Fp[MFC5_M2_FPS ? 0 : v];
                    if (DIR == 0) r[v] = inv_ds * (fp - F[v]);
                    else r[v] = rin[v] + inv_ds * (fp - F[v]);
                }
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    if (alpha_iface) r[S::IA + f] = r[S::IA + f] + inv_ds * acoef[f] * (Fsp[f] - Fs[f]);
                    else r[S::IA + f] = r[S::IA + f] + inv_ds * acoef[f] * (Fs[0] - Fsp[0]);
                }
                if (rk) {
#pragma unroll
                    for (int v = 0; v < N; v++) r[v] = div_by_const(P.rk0 * q1in[v] + P.rk1 * q2in[v] + P.rk2 * P.dt * r[v], P.rk3, P.irk3);
                    if (P.bf_on) apply_body_forces<NF, ND>(P, r);
#if MFC5_M2_DIET
                    {
                        double* po = P.qout + c;
#pragma unroll
                        for (int v = 0; v < N; v++) { *po = r[v]; po += fs; }
                    }
                    if (P.prim_next) {
                        double pr[N];
                        cons2prim_cell_fast<NF, ND>(P, r, pr);
                        double* pn = P.prim_next + c;
#pragma unroll
                        for (int v = 0; v < N; v++) { *pn = pr[v]; pn += fs; }
                    }
                } else {
                    double* po = P.rhs + c;
#pragma unroll
                    for (int v = 0; v < N; v++) { *po = r[v]; po += fs; }
                }
#else
#pragma unroll
                    for (int v = 0; v < N; v++) P.qout[(ix)((ix)v * fs + c)] = r[v];
                    if (P.prim_next) {   // primitive variables of the new state for the next stage (over the consumed rhs entry)
                        double pr[N];
                        cons2prim_cell_fast<NF, ND>(P, r, pr);
#pragma unroll
                        for (int v = 0; v < N; v++) P.prim_next[(ix)((ix)v * fs + c)] = pr[v];
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < N; v++) P.rhs[(ix)((ix)v * fs + c)] = r[v];
                }
#endif
            }
#pragma unroll
            for (int v = 0; v < N; v++) {
                if (MFC5_M2_FPS) sFp[v * NT + tid] = F[v];
                else Fp[MFC5_M2_FPS ? 0 : v] = F[v];
            }
#pragma unroll
            for (int f = 0; f < NF; f++) Fsp[f] = Fs[f];
        }
        tail = tail + 1;
        if (tail >= NSLOT) tail = 0;
    }
}

// -------------------------------------------------------------------------------------------------
// Reductions: ICFL max (m_data_output.fpp:159-224, m_sim_helpers.fpp:45-130), dt min
// (m_sim_helpers.fpp:182-249) and NaN count.  Warp shuffle reduction + one atomic per warp.
// out[0] = max icfl (as ordered-uint bits), out[1] = min dt_local bits, out[2] = nan count.
// -------------------------------------------------------------------------------------------------
template <int NF, int ND>
__global__ void __launch_bounds__(256) k_reduce(const __grid_constant__ DevP P, double dt, unsigned long long* out) {
    using S = Shape<NF, ND>;
    const int nx = P.s[0] + 1, ny = P.s[1] + 1, nz = P.s[2] + 1;
    const long long tot = (long long)nx * ny * nz;
    double icfl_max = 0.0, dtl_min = 1.0e300;
    unsigned long long nans = 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx);
        const long long r = t / nx;
        const int j = (int)(r % ny), k = (int)(r / ny);
        const long long c = P.base + i + j * P.sy + k * P.sz;
        double arho[NF], alf[NF], rho, gam, pinf, qv, vel[3] = {0, 0, 0}, vs = 0;
#pragma unroll
        for (int f = 0; f < NF; f++) { arho[f] = P.prim[f * P.fs + c]; alf[f] = P.prim[(S::IA + f) * P.fs + c]; }
        mixture<NF>(P, arho, alf, 2, rho, gam, pinf, qv);
#pragma unroll
        for (int a = 0; a < ND; a++) { vel[a] = P.prim[(S::IM + a) * P.fs + c]; vs = vs + vel[a] * vel[a]; }
        const double pres = P.prim[S::IE * P.fs + c];
        const double E = gam * pres + pinf + 5.e-1 * rho * vs + qv;
        const double H = (E + pres) / rho;
        const double cs = sound_speed(P, pres, rho, gam, H, alf, vs, qv);
        double dtl = P.T[0].ds[i + P.b] / (fabs(vel[0]) + cs);
        if (ND > 1) dtl = fmin(dtl, P.T[1].ds[j + P.b] / (fabs(vel[1]) + cs));
        if (ND > 2) dtl = fmin(dtl, P.T[2].ds[k + P.b] / (fabs(vel[2]) + cs));
        double icfl;
        if (ND == 1) icfl = (dt / P.T[0].ds[i + P.b]) * (fabs(vel[0]) + cs);
        else icfl = dt / dtl;
        if (P.viscous) {   // viscous limit of s_compute_dt_from_cfl (m_sim_helpers.fpp:214-227) with the mixture Reynolds numbers of
            // s_convert_species_to_mixture_variables_kernel (m_variables_conversion.fpp:256-268; alpha after the limiter)
            double Re[2];
            interface_reynolds<NF>(P, alf, Re);
            double dsm = P.T[0].ds[i + P.b];
            if (ND > 1) dsm = fmin(dsm, P.T[1].ds[j + P.b]);
            if (ND > 2) dsm = fmin(dsm, P.T[2].ds[k + P.b]);
            const double inv0 = (ND == 2) ? (1.0 / Re[0]) / rho : 1.0 / (rho * Re[0]);
            const double inv1 = (ND == 2) ? (1.0 / Re[1]) / rho : 1.0 / (rho * Re[1]);
            dtl = fmin(dtl, (dsm * dsm) / fmax(inv0, inv1));
        }
        if (icfl != icfl) nans++;
        icfl_max = fmax(icfl_max, icfl);
        dtl_min = fmin(dtl_min, dtl);
#pragma unroll
        for (int v = 0; v < S::N; v++) { const double x = P.q1[v * P.fs + c]; if (x != x) nans++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        icfl_max = fmax(icfl_max, __shfl_xor_sync(0xffffffffu, icfl_max, o));
        dtl_min = fmin(dtl_min, __shfl_xor_sync(0xffffffffu, dtl_min, o));
        nans += __shfl_xor_sync(0xffffffffu, nans, o);
    }
    if ((threadIdx.x & 31) == 0) {
        // positive doubles order like their bit patterns
        atomicMax(out + 0, (unsigned long long)__double_as_longlong(icfl_max));
        atomicMin(out + 1, (unsigned long long)__double_as_longlong(dtl_min));
        if (nans) atomicAdd(out + 2, nans);
    }
}

// packed halo faces for processor boundaries (replaces the pack/unpack loops of
// s_mpi_sendrecv_variables_buffers, src/common/m_mpi_common.fpp:620-804,839-1065), every processor face
// of the block in ONE launch (blockIdx.y = 2 * direction + side).  Only the transverse interior travels
// (the reference also ships the edge/corner ghosts, which the inviscid path never reads).
struct HaloBufs {
    double* buf[6];
};
#ifdef MFC5_HOST_TU
// full = 1 (viscous path): one direction per launch (face0 = 2 d), transverse ranges of the directions BEFORE d widened by the
// ghost layers, which those directions' fills / exchanges have already written -- edges and corners travel as in the reference's
// x -> y -> z sequence (m_mpi_common.fpp:620-1065 called per direction from s_populate_variables_buffers)
__global__ void __launch_bounds__(256) k_halo_faces(const __grid_constant__ DevP P, const __grid_constant__ HaloBufs B, int pack, int full = 0, int face0 = 0) {
    const int face = blockIdx.y + face0;
    const int d = face >> 1, side = face & 1;
    if (d >= P.nd || P.bc[d][side] < 0) return;
    int lo[3] = {0, 0, 0}, hi[3] = {P.s[0], P.s[1], P.s[2]};
    if (full) for (int e = 0; e < d; e++) { lo[e] = -P.b; hi[e] = P.s[e] + P.b; }
    // pack: the b interior layers next to the face; unpack: the b ghost layers beyond it
    const int first = pack ? (side == 0 ? 0 : P.s[d] - P.b + 1) : (side == 0 ? -P.b : P.s[d] + 1);
    lo[d] = first; hi[d] = first + P.b - 1;
    const int n0 = hi[0] - lo[0] + 1, n1 = hi[1] - lo[1] + 1, n2 = hi[2] - lo[2] + 1;
    const long long per = (long long)n0 * n1 * n2, tot = per * P.N;
    double* buf = B.buf[face];
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t / per);
        long long r = t % per;
        const int i = (int)(r % n0) + lo[0]; r /= n0;
        const int j = (int)(r % n1) + lo[1];
        const int k = (int)(r / n1) + lo[2];
        double* f = P.primw + v * P.fs + P.base + i + j * P.sy + k * P.sz;
        if (pack) buf[t] = *f; else *f = buf[t];
    }
}

// Host <-> device field copies go through a contiguous staging image (one large DMA per field at full PCIe
// rate instead of a pitched 3-D copy of 1.6 kB rows); this kernel moves the copy region between the
// staging image [N][hz][hy][hx] and the padded device layout.  to_dev = 1: staging -> device.
struct RepackGeom {
    int N, wx, wy, wz;          // copy region
    long long dev_off, sy, sz, fs;      // device: offset of the region's first cell inside a field, strides, field size
    long long st_off, hx, hxy, hfs;     // staging: the same for the host-shaped image
    int gx, gy, gz, zero_ghosts;        // download: region cells outside [g, w - g) per direction are written as zeros
};
__global__ void __launch_bounds__(256) k_repack(double* __restrict__ dev, double* __restrict__ stage, const __grid_constant__ RepackGeom G, int to_dev) {
    const long long per = (long long)G.wx * G.wy * G.wz, tot = per * G.N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t / per);
        long long r = t % per;
        const int i = (int)(r % G.wx); r /= G.wx;
        const int j = (int)(r % G.wy);
        const int k = (int)(r / G.wy);
        double* d = dev + v * G.fs + G.dev_off + i + j * G.sy + k * G.sz;
        double* s = stage + v * G.hfs + G.st_off + i + j * G.hx + k * G.hxy;
        if (to_dev) *d = *s;
        else if (G.zero_ghosts && (i < G.gx || i >= G.wx - G.gx || j < G.gy || j >= G.wy - G.gy || k < G.gz || k >= G.wz - G.gz)) *s = 0.0;
        else *s = *d;
    }
}
#endif  // MFC5_HOST_TU

}  // namespace mfc5
