// mfc5.cu -- C ABI of libmfc5.so (include/mfc5.h): device state, table set-up, stage orchestration.
//
// One RK stage (reference s_compute_rhs m_rhs.fpp:574-902 + RK update m_time_steppers.fpp:486-507):
//   k_cons2prim -> [k_bc_fill | halo exchange] per direction -> per direction { k_cbc ; fused sweep }
// with the RK update fused into the last sweep.  There is no CPU fallback: every entry point fails
// loudly if CUDA is unavailable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler is attached

#include "../../include/mfc5.h"
#define MFC5_HOST_TU
#include "mfc5_host.cuh"

using namespace mfc5;

// NVTX ranges with the reference's names (m_time_steppers.fpp:453 TIMESTEP, m_rhs.fpp:594 COMPUTE-RHS, :640 RHS-CONVERT,
// :635/:644 RHS-COMMUNICATION, :912 RHS-RECONSTRUCTION, :1030 RHS-RIEMANN-SOLVER, :1041 RHS-ADVECTION-SRC -- the last three
// are one fused kernel per direction here, so they share one range).  Host-side enqueue ranges: under CUDA-graph replay a
// whole step is one launch inside TIMESTEP.
struct NvtxRange {
    bool open = true;
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    void end() { if (open) { nvtxRangePop(); open = false; } }
    ~NvtxRange() { end(); }
};

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

// ---- NCCL, loaded lazily so single-GPU use has no dependency on it --------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail("cannot dlopen libnccl.so.2");
#define SYM(f, name) *(void**)(&g_nccl.f) = dlsym(g_nccl.lib, name); if (!g_nccl.f) return fail("missing NCCL symbol " name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}
#define NK(call)                                                                                          \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return fail(std::string(#call) + ": " + g_nccl.GetErrorString(r_));                  \
    } while (0)
enum { NCCL_FLOAT64 = 8, NCCL_UINT64 = 5, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };

// ---- handle ---------------------------------------------------------------------------------------
struct mfc5_handle {
    mfc5_params P;
    int nd, nf, N, b, pn, ns;
    int s[3];
    int ox, oy, oz;
    long long px, ey, ez, fs;
    double* d_q[2] = {nullptr, nullptr};
    // two primitive-variable buffers: the RK-fused last sweep of a stage writes the primitives of the new state into the
    // other one, so that the next stage needs no cons -> prim pass (prim_valid: d_prim[pp] matches the current state)
    double* d_prim[2] = {nullptr, nullptr};
    double* d_rhs = nullptr;
    // case-optimised path: the rhs buffer doubles as the second primitive buffer (d_prim[1] == d_rhs): the RK-fused last sweep
    // stores the primitive variables of the new state over the rhs entries it has consumed, and the two buffers swap roles
    bool alias = false;
    int pp = 0;
    bool prim_valid = false;
    double mytime = 0.0;   // simulation time of the state (body forces); advanced by dt after every step
    int fuse_c2p = 0;
    int cur = 0;
    // host grid (ghosted): scb[d][i + 1 + b], ds[d][i + b]
    std::vector<double> scb[3], ds[3], scc[3];   // (scc: cell centres, [i + b])
    bool grid_set[3] = {false, false, false};
    bool setup = false;
    double* d_tab[3][10] = {};  // cL cR dL dR bt ds + combined rows + their transpose + x-segment layout + cell centres
    int uniform[3] = {1, 1, 1};
    CbcCoef cbc[3][2];
    double* d_cbcbuf[3][2] = {};
    double* d_dirich[3][2] = {};
    long long plane[3];
    cudaStream_t st = nullptr, st_comm = nullptr, st_bc = nullptr;
    cudaEvent_t ev_bc = nullptr;
    int use_side_stream = 1;
    cudaEvent_t ev_stage[4] = {}, ev_comm = nullptr, ev_comp = nullptr;
    // per-kernel-class device timing of the last mfc5_tvd_rk call (events on the compute stream)
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> seg_cls;  // class of segment i = events (2i, 2i+1)
    double cls_ms[MFC5_NUM_KCLASS] = {};
    long long cls_launches[MFC5_NUM_KCLASS] = {};
    bool timing = false;
    unsigned long long* d_red = nullptr;
    unsigned long long* d_red_all = nullptr;   // [nranks][4]: every rank's partial results (one ncclAllGather per reduction)
    long long launches = 0;
    DevP D;
    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    double* d_send[3][2] = {};
    double* d_recv[3][2] = {};
    bool any_phys = false, any_nb = false, any_cbc = false;
    CbcCoef* d_cbccoef = nullptr;  // [6], index 2*dir + side
    cudaEvent_t ev_prim = nullptr, ev_halo = nullptr;
    // graph replay of a whole step, keyed by (cur, dt)
    cudaGraphExec_t gexec[4] = {nullptr, nullptr, nullptr, nullptr};  // key = 2 * cur + pp
    double gdt[4] = {0, 0, 0, 0};
    long long glaunches[4] = {0, 0, 0, 0};  // kernel launches inside each captured step
    int gcur1[4] = {0, 0, 0, 0}, gpp1[4] = {0, 0, 0, 0};  // buffer parities after the captured step
    long long launches_captured = 0;
    bool warmed = false;              // one step has run uncaptured (lazy allocations, function attributes)
    int use_graph = 1;
    bool graph_fallback = false;      // a capture failed and the handle fell back to eager launches
};

static double* rhs_ptr(const mfc5_handle* h) { return h->alias ? h->d_prim[h->pp ^ 1] : h->d_rhs; }

static bool is_cbc(int bc) { return bc <= MFC5_BC_CHAR_SLIP_WALL && bc >= MFC5_BC_CHAR_SUP_OUTFLOW; }

extern "C" const char* mfc5_last_error(void) { return g_err.c_str(); }
extern "C" const char* mfc5_version(void) { return "mfc5-b200 0.1 (sm_100a, fp64)"; }
extern "C" int mfc5_sys_size(const mfc5_handle* h) { return h->N; }
extern "C" int mfc5_buff_size(const mfc5_handle* h) { return h->b; }
extern "C" long long mfc5_kernel_launches(const mfc5_handle* h) { return h->launches; }
extern "C" double mfc5_last_kernel_ms(const mfc5_handle* h, int kclass, long long* launches) {
    if (!h || kclass < 0 || kclass >= MFC5_NUM_KCLASS) return -1.0;
    if (launches) *launches = h->cls_launches[kclass];
    return h->cls_ms[kclass];
}
extern "C" void* mfc5_compute_stream(const mfc5_handle* h) { return (void*)h->st; }

extern "C" int mfc5_create(const mfc5_params* p, mfc5_handle** out) {
    if (!p || !out) return fail("null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("no CUDA device available (libmfc5 has no CPU fallback): ") + cudaGetErrorString(e));
    if (p->model_eqns != 2) return fail("only model_eqns = 2 (5-equation model) is supported");
    if (p->num_fluids < 1 || p->num_fluids > 3) return fail("num_fluids must be 1..3");
    mfc5_params pe = *p;   // effective parameters: MUSCL borrows the stencil / buffer geometry of the WENO of the same width
    if (p->muscl_order != 0) {
        if (p->muscl_order < 1 || p->muscl_order > 2 || p->muscl_lim < 0 || p->muscl_lim > 5) return fail("muscl_order must be 1 or 2, muscl_lim 0..5");
        pe.weno_order = 2 * p->muscl_order + 1;   // buff_size = muscl_order + 2 (m_helper_basic.fpp:134-135)
    }
    p = &pe;
    if (p->weno_order != 1 && p->weno_order != 3 && p->weno_order != 5 && p->weno_order != 7) return fail("weno_order must be 1, 3, 5 or 7");
    if (p->weno_order == 7 && p->wenoz && !(p->wenoz_q == 2.0 || p->wenoz_q == 3.0 || p->wenoz_q == 4.0)) return fail("wenoz at weno_order = 7 needs wenoz_q = 2, 3 or 4");
    if (p->weno_order == 7 && (p->int_comp != 0 || p->null_weights)) return fail("weno_order = 7: int_comp and null_weights are not available");
    if (p->riemann_solver != 1 && p->riemann_solver != 2 && p->riemann_solver != 5)
        return fail("riemann_solver must be 1 (HLL), 2 (HLLC) or 5 (Lax-Friedrichs)");
    if (p->mapped_weno && p->wenoz) return fail("mapped_weno and wenoz are mutually exclusive");
    if (p->int_comp != 0 && (p->int_comp != 1 || p->num_fluids != 2 || (p->muscl_order > 0 ? p->muscl_order : p->weno_order) == 1))
        return fail("int_comp: only THINC (1), with num_fluids = 2 and a reconstruction order above 1");
    if (p->mp_weno && (p->weno_order != 5 || p->muscl_order != 0)) return fail("mp_weno requires weno_order = 5");
    if (p->teno && (p->mapped_weno || p->wenoz || (p->weno_order != 5 && p->weno_order != 7) || p->muscl_order != 0 || !(p->teno_CT > 0.0)))
        return fail("teno needs weno_order = 5 or 7, teno_CT > 0 and neither mapped_weno nor wenoz");
    if (p->alt_soundspeed && (p->num_fluids != 2 || p->avg_state != 2 || (p->riemann_solver == 1 && !p->hll_u_interface)))
        return fail("alt_soundspeed needs 2 fluids, avg_state = 2 and a u-interface solver (HLLC, or HLL with hll_u_interface)");
    if (p->low_Mach < 0 || p->low_Mach > 2 || (p->low_Mach == 2 && p->riemann_solver != 2)) return fail("low_Mach must be 0, 1 or 2 (2: HLLC only)");
    if (p->time_stepper < 1 || p->time_stepper > 3) return fail("time_stepper must be 1..3");
    if (p->m < 0 || p->n < 0 || p->p < 0 || (p->n == 0 && p->p > 0)) return fail("invalid m, n, p");
    // m_checker.fpp:52-69: at least num_stcls_min*weno_order cells per direction
    const int recon_order = p->muscl_order > 0 ? p->muscl_order : p->weno_order;
    const int need = 5 * recon_order;   // num_stcls_min * weno_order for every order, 1 included (s_check_inputs_weno)
    const int ss[3] = {p->m, p->n, p->p};
    for (int d = 0; d < 3; d++)
        if ((d == 0 || ss[d] > 0) && ss[d] + 1 < need) return fail("too few cells for the WENO stencil (m_checker.fpp:52-69)");
    mfc5_handle* h = new mfc5_handle();
    h->P = *p;
    h->s[0] = p->m; h->s[1] = p->n; h->s[2] = p->p;
    h->nd = 1 + (p->n > 0) + (p->p > 0);
    h->nf = p->num_fluids;
    h->N = 2 * h->nf + h->nd + 1;
    h->pn = (p->weno_order - 1) / 2;
    h->ns = h->pn + 1;
    h->b = h->pn + 2;  // m_helper_basic.fpp:127-132
    if (p->viscous && p->muscl_order == 0) h->b = 2 * h->pn + 2;   // viscous WENO: 2 weno_polyn + 2
    if (p->viscous) {
        const char* why = nullptr;
        if (p->muscl_order > 0) why = "viscous: not available with MUSCL reconstruction";
        else if (p->riemann_solver == 5) why = "viscous: not available with the Lax-Friedrichs solver";
        else if (p->weno_Re_flux) why = "viscous: weno_Re_flux = T is not available (finite-difference gradients, weno_Re_flux = F, are)";
        else if (p->weno_order == 7) why = "viscous: not available at weno_order = 7";
        bool any = false;
        for (int f = 0; f < p->num_fluids; f++) any = any || p->fluid_Re[f][0] > 0 || p->fluid_Re[f][1] > 0;
        if (!why && !any) why = "viscous = T needs at least one fluid_pp(i)%Re(1) or %Re(2)";
        if (why) { delete h; return fail(why); }
    }
    for (int d = 0; d < h->nd; d++)
        for (int sd = 0; sd < 2; sd++)
            if (is_cbc(p->bc[d][sd]) && (p->weno_order == 1 || p->weno_order == 7 || p->muscl_order > 0)) {
                delete h;
                return fail("characteristic BCs need weno_order 3 or 5 (and are not available with MUSCL)");
            }
    CK(cudaSetDevice(p->device));
    // HBM layout: SoA, x contiguous; interior x = 0 sits on a 128-byte boundary
    h->ox = 16; h->oy = h->nd > 1 ? h->b : 0; h->oz = h->nd > 2 ? h->b : 0;
    h->px = ((h->ox + h->s[0] + 1 + h->b + 15) / 16) * 16;
    h->ey = h->s[1] + 1 + 2 * h->oy;
    h->ez = h->s[2] + 1 + 2 * h->oz;
    h->fs = h->px * h->ey * h->ez;
    const size_t nb = (size_t)h->N * h->fs * sizeof(double);
    CK(cudaMalloc(&h->d_q[0], nb)); CK(cudaMalloc(&h->d_q[1], nb));
    CK(cudaMalloc(&h->d_prim[0], nb)); CK(cudaMalloc(&h->d_rhs, nb));  // d_prim[1]: only with MFC5_FUSE_C2P (finalize_setup)
    CK(cudaMemset(h->d_q[0], 0, nb)); CK(cudaMemset(h->d_q[1], 0, nb));
    CK(cudaMemset(h->d_prim[0], 0, nb)); CK(cudaMemset(h->d_rhs, 0, nb));
    CK(cudaMalloc(&h->d_red, 4 * sizeof(unsigned long long)));
    int lo, hi;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&h->st_comm, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&h->st_bc, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&h->ev_bc, cudaEventDisableTiming));
    for (auto& ev : h->ev_stage) CK(cudaEventCreate(&ev));
    CK(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_comp, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_prim, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    CK(cudaMalloc(&h->d_cbccoef, 6 * sizeof(CbcCoef)));
    *out = h;
    return 0;
}

extern "C" int mfc5_destroy(mfc5_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->P.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) cudaFree(h->d_q[i]);
    for (int i = 0; i < 4; i++) if (h->gexec[i]) cudaGraphExecDestroy(h->gexec[i]);
    cudaFree(h->d_prim[0]); if (!h->alias) cudaFree(h->d_prim[1]); cudaFree(h->d_rhs); cudaFree(h->d_red); cudaFree(h->d_red_all);
    for (int d = 0; d < 3; d++) {
        for (int k = 0; k < 10; k++) cudaFree(h->d_tab[d][k]);
        for (int k = 0; k < 2; k++) { cudaFree(h->d_cbcbuf[d][k]); cudaFree(h->d_dirich[d][k]); cudaFree(h->d_send[d][k]); cudaFree(h->d_recv[d][k]); }
    }
    for (auto& ev : h->ev_stage) cudaEventDestroy(ev);
    for (auto& ev : h->ev_pool) cudaEventDestroy(ev);
    cudaEventDestroy(h->ev_comm); cudaEventDestroy(h->ev_comp);
    cudaEventDestroy(h->ev_prim); cudaEventDestroy(h->ev_halo);
    cudaFree(h->d_cbccoef);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    cudaEventDestroy(h->ev_bc);
    cudaStreamDestroy(h->st); cudaStreamDestroy(h->st_comm); cudaStreamDestroy(h->st_bc);
    delete h;
    return 0;
}

// ---- grid ghost metrics: m_boundary_common.fpp:404-466, m_mpi_common.fpp:1745-1796 ------------------
extern "C" int mfc5_set_grid(mfc5_handle* h, int d, const double* s_cb, const double* gbeg, const double* gend) {
    if (!h || d < 0 || d >= h->nd || !s_cb) return fail("mfc5_set_grid: bad argument");
    const int s = h->s[d], b = h->b;
    h->scb[d].assign(s + 2 + 2 * b, 0.0);
    h->ds[d].assign(s + 1 + 2 * b, 0.0);
    double* cb = h->scb[d].data() + 1 + b;
    double* w = h->ds[d].data() + b;
    for (int i = -1; i <= s; i++) cb[i] = s_cb[i + 1];
    for (int i = 0; i <= s; i++) w[i] = cb[i] - cb[i - 1];  // m_start_up.fpp:190
    const int bcb = h->P.bc[d][0], bce = h->P.bc[d][1];
    if ((bcb >= 0 && !gbeg) || (bce >= 0 && !gend)) return fail("mfc5_set_grid: neighbour cell widths required on processor boundaries");
    for (int i = 1; i <= b; i++) {
        if (bcb >= 0) w[-i] = gbeg[b - i];
        else w[-i] = w[bcb == MFC5_BC_PERIODIC ? s - i + 1 : bcb == MFC5_BC_REFLECTIVE ? i - 1 : 0];
        if (bce >= 0) w[s + i] = gend[i - 1];
        else w[s + i] = w[bce == MFC5_BC_PERIODIC ? i - 1 : bce == MFC5_BC_REFLECTIVE ? s - i + 1 : s];
    }
    for (int i = 1; i <= b; i++) cb[-1 - i] = cb[-i] - w[-i];
    for (int i = 1; i <= b; i++) cb[s + i] = cb[s + i - 1] + w[s + i];
    // cell centres: m_start_up.fpp:191 and the ghost recurrence of s_populate_grid_bc_direction (m_boundary_common.fpp:449-465)
    h->scc[d].assign(s + 1 + 2 * b, 0.0);
    double* cc = h->scc[d].data() + b;
    for (int i = 0; i <= s; i++) cc[i] = cb[i - 1] + w[i] / 2.0;
    for (int i = 1; i <= b; i++) cc[-i] = cc[1 - i] - (w[1 - i] + w[-i]) / 2.0;
    for (int i = 1; i <= b; i++) cc[s + i] = cc[s + i - 1] + (w[s + i - 1] + w[s + i]) / 2.0;
    h->grid_set[d] = true;
    h->setup = false;
    return 0;
}

// ---- WENO tables of order 7 (k = 4): s_compute_weno_coefficients m_weno.fpp:362-851 -------------------------------------
// The reference lists these as ~500 lines of closed forms; here they come from the definitions those forms expand
// (Shu 1997, eq. 2.20: the candidate polynomial of a k-cell stencil is the derivative of the Lagrange interpolant of the
// primitive function through its k + 1 cell boundaries, written on the cell-value differences dvd; ideal weights: the
// combination that reproduces the (2k-1)-cell polynomial at the cell boundary; smoothness indicators: Jiang & Shu,
// beta = sum_l dx^(2l-1) int (d^l p / dx^l)^2 dx as a quadratic form in the dvd's).  Same operations in the same order as
// the CPU oracle's construction, which is pinned by the reference's WENO7 goldens, so the two tables are bit-identical.
namespace wgen {
constexpr int KM = 4;
struct Poly { double c[8]; int n; };
static Poly mul_lin(const Poly& a, double r, double scale) {   // a * (x - r) * scale
    Poly o; o.n = a.n + 1;
    for (int i = 0; i < o.n; i++) o.c[i] = 0.0;
    for (int i = 0; i < a.n; i++) { o.c[i + 1] += a.c[i] * scale; o.c[i] -= a.c[i] * r * scale; }
    return o;
}
static Poly deriv(const Poly& a) {
    Poly o; o.n = a.n > 1 ? a.n - 1 : 1; o.c[0] = 0.0;
    for (int i = 1; i < a.n; i++) o.c[i - 1] = a.c[i] * (double)i;
    return o;
}
static Poly axpy(const Poly& y, double a, const Poly& x) {      // y + a x
    Poly o; o.n = y.n > x.n ? y.n : x.n;
    for (int i = 0; i < o.n; i++) o.c[i] = (i < y.n ? y.c[i] : 0.0) + a * (i < x.n ? x.c[i] : 0.0);
    return o;
}
static double eval(const Poly& a, double x) { double v = 0.0; for (int i = a.n - 1; i >= 0; i--) v = v * x + a.c[i]; return v; }
static double int_prod(const Poly& a, const Poly& b, double L) {   // int_0^L a b dx
    double prod[16]; const int n = a.n + b.n - 1;
    for (int i = 0; i < n; i++) prod[i] = 0.0;
    for (int i = 0; i < a.n; i++) for (int j = 0; j < b.n; j++) prod[i + j] += a.c[i] * b.c[j];
    double v = 0.0, Lp = L;
    for (int i = 0; i < n; i++) { v += prod[i] * Lp / (double)(i + 1); Lp *= L; }
    return v;
}
// c_j(x), j = 0..kk-1: weights of the cell averages in the reconstruction polynomial over the cells bounded by X[0..kk]
static void cell_weights(const double* X, int kk, Poly* cj) {
    Poly dL[2 * KM];
    for (int m = 0; m <= kk; m++) {
        Poly l; l.n = 1; l.c[0] = 1.0;
        for (int n = 0; n <= kk; n++) if (n != m) l = mul_lin(l, X[n], 1.0 / (X[m] - X[n]));
        dL[m] = deriv(l);
    }
    for (int j = 0; j < kk; j++) {
        Poly acc; acc.n = 1; acc.c[0] = 0.0;
        for (int m = j + 1; m <= kk; m++) acc = axpy(acc, 1.0, dL[m]);
        for (int i = 0; i < acc.n; i++) acc.c[i] *= (X[j + 1] - X[j]);
        cj[j] = acc;
    }
}
// tables of cell i for order 2k-1; xb = the 2k cell boundaries around it.  cL / cR[sidx][r] (r = 0: highest dvd; sidx = 0: the
// rightmost stencil), dL / dR[sidx], bt[sidx][0 .. k(k-1)/2 - 1]
static void tables(int k, const double* xb, double cLo[][KM - 1], double cRo[][KM - 1], double* dLo, double* dRo, double bto[][6]) {
    double X[2 * KM];
    const double x0 = xb[k - 1];
    for (int m = 0; m < 2 * k; m++) X[m] = xb[m] - x0;
    const double dxi = X[k] - X[k - 1];
    Poly cw[KM][KM];
    for (int l = 0; l < k; l++) {
        cell_weights(X + (k - 1 - l), k, cw[l]);
        Poly a[KM - 1];
        for (int e = 0; e < k - 1; e++) {
            Poly acc; acc.n = 1; acc.c[0] = 0.0;
            if (e >= l) { for (int j = e + 1; j < k; j++) acc = axpy(acc, 1.0, cw[l][j]); }
            else        { for (int j = 0; j <= e; j++) acc = axpy(acc, -1.0, cw[l][j]); }
            a[e] = acc;
        }
        for (int r = 0; r < k - 1; r++) {
            const int e = k - 2 - r;
            cLo[l][r] = eval(a[e], 0.0);
            cRo[l][r] = eval(a[e], dxi);
        }
        double B[KM - 1][KM - 1];
        for (int e = 0; e < k - 1; e++) for (int f = 0; f < k - 1; f++) B[e][f] = 0.0;
        Poly da[KM - 1];
        for (int e = 0; e < k - 1; e++) da[e] = a[e];
        double w = dxi;
        for (int ord = 1; ord <= k - 1; ord++) {
            for (int e = 0; e < k - 1; e++) da[e] = deriv(da[e]);
            for (int e = 0; e < k - 1; e++) for (int f = 0; f < k - 1; f++) B[e][f] += w * int_prod(da[e], da[f], dxi);
            w *= dxi * dxi;
        }
        int n = 0;
        for (int r = 0; r < k - 1; r++) for (int r2 = r; r2 < k - 1; r2++) {
            const int e = k - 2 - r, f = k - 2 - r2;
            bto[l][n++] = (r == r2) ? B[e][e] : 2.0 * B[e][f];
        }
    }
    Poly cbig[2 * KM];
    cell_weights(X, 2 * k - 1, cbig);
    for (int side = 0; side < 2; side++) {
        const double xe = side == 0 ? 0.0 : dxi;
        double* dd = side == 0 ? dLo : dRo;
        for (int l = k - 1; l >= 0; l--) {
            const int g = k - 1 - l;
            double rhs = eval(cbig[g], xe);
            for (int l2 = k - 1; l2 > l; l2--) rhs -= dd[l2] * eval(cw[l2][g - (k - 1 - l2)], xe);
            dd[l] = rhs / eval(cw[l][0], xe);
        }
    }
}
}  // namespace wgen

// ---- WENO tables: s_compute_weno_coefficients, m_weno.fpp:196-215 (order 3), :233-341 (order 5) -----
static void weno_tables(const mfc5_handle* h, int d, std::vector<double> (&tab)[5], int& uniform) {
    const int s = h->s[d], b = h->b, pn = h->pn, ns = h->ns;
    const int nslot = s + 1 + 2 * b;
    const int pr = pn > 0 ? pn : 1;
    tab[0].assign((size_t)nslot * ns * pr, 0.0);  // cL
    tab[1].assign((size_t)nslot * ns * pr, 0.0);  // cR
    tab[2].assign((size_t)nslot * ns, 0.0);       // dL
    tab[3].assign((size_t)nslot * ns, 0.0);       // dR
    tab[4].assign((size_t)nslot * ns * (pn == 3 ? 6 : 3), 0.0);   // beta (order 7: 6 entries per stencil)
    const double* x = h->scb[d].data() + 1 + b;
    auto q2 = [](double v) { return v * v; };
    auto CL = [&](int j, int q, int r) -> double& { return tab[0][((size_t)(j + b) * ns + q) * pr + r]; };
    auto CR = [&](int j, int q, int r) -> double& { return tab[1][((size_t)(j + b) * ns + q) * pr + r]; };
    auto DL = [&](int q, int j) -> double& { return tab[2][(size_t)(j + b) * ns + q]; };
    auto DR = [&](int q, int j) -> double& { return tab[3][(size_t)(j + b) * ns + q]; };
    auto BT = [&](int j, int q, int r) -> double& { return tab[4][((size_t)(j + b) * ns + q) * 3 + r]; };
    const int bcb = h->P.bc[d][0], bce = h->P.bc[d][1];
    if (h->P.weno_order == 3) {
        for (int i = -b - 1 + pn; i <= s + b - 1 - pn; i++) {
            const int j = i + 1;
            const double xm = x[i - 1], x0 = x[i], x1 = x[i + 1], x2 = x[i + 2];
            CR(j, 0, 0) = (x0 - x1) / (x0 - x2);
            CR(j, 1, 0) = (x0 - x1) / (xm - x1);
            CL(j, 0, 0) = -CR(j, 0, 0);
            CL(j, 1, 0) = -CR(j, 1, 0);
            DR(0, j) = (xm - x1) / (xm - x2);
            DL(0, j) = (xm - x0) / (xm - x2);
            DR(1, j) = 1.0 - DR(0, j);
            DL(1, j) = 1.0 - DL(0, j);
            BT(j, 0, 0) = 4.0 * q2(x0 - x1) / q2(x0 - x2);
            BT(j, 1, 0) = 4.0 * q2(x0 - x1) / q2(xm - x1);
        }
        if (h->P.null_weights) {  // m_weno.fpp:219-229
            if (bcb == MFC5_BC_RIEMANN_EXTRAP) { DR(1, 0) = 0; DR(0, 0) = 1; DL(1, 0) = 0; DL(0, 0) = 1; }
            if (bce == MFC5_BC_RIEMANN_EXTRAP) { DR(0, s) = 0; DR(1, s) = 1; DL(0, s) = 0; DL(1, s) = 1; }
        }
    } else if (h->P.weno_order == 5) {
        for (int i = -b - 1 + pn; i <= s + b - 1 - pn; i++) {
            const int j = i + 1;
            const double a = x[i - 2], bb = x[i - 1], c = x[i], e = x[i + 1], f = x[i + 2], g = x[i + 3];
            // a = s_cb(i-2), bb = s_cb(i-1), c = s_cb(i), e = s_cb(i+1), f = s_cb(i+2), g = s_cb(i+3)
            CR(j, 0, 0) = ((c - e) * (e - f)) / ((c - g) * (g - e));
            CR(j, 1, 0) = ((bb - e) * (e - c)) / ((bb - f) * (f - c));
            CR(j, 1, 1) = ((c - e) * (e - f)) / ((bb - e) * (bb - f));
            CR(j, 2, 1) = ((c - e) * (e - bb)) / ((a - c) * (a - e));
            CL(j, 0, 0) = ((e - c) * (c - f)) / ((c - g) * (g - e));
            CL(j, 1, 0) = ((c - bb) * (c - e)) / ((bb - f) * (c - f));
            CL(j, 1, 1) = ((e - c) * (c - f)) / ((bb - e) * (bb - f));
            CL(j, 2, 1) = ((bb - c) * (c - e)) / ((a - c) * (a - e));
            CR(j, 0, 1) = ((c - f) + (e - g)) / ((c - f) * (c - g)) * ((c - e));
            CR(j, 2, 0) = ((a - e) + (bb - e)) / ((bb - e) * (e - a)) * ((e - c));
            CL(j, 0, 1) = ((c - f) + (c - g)) / ((c - f) * (c - g)) * ((e - c));
            CL(j, 2, 0) = ((a - c) + (bb - e)) / ((a - e) * (e - bb)) * ((c - e));
            DR(0, j) = ((a - e) * (e - bb)) / ((a - g) * (g - bb));
            DR(2, j) = ((e - f) * (e - g)) / ((a - f) * (a - g));
            DL(0, j) = ((a - c) * (c - bb)) / ((a - g) * (g - bb));
            DL(2, j) = ((c - f) * (c - g)) / ((a - f) * (a - g));
            DR(1, j) = 1.0 - DR(0, j) - DR(2, j);
            DL(1, j) = 1.0 - DL(0, j) - DL(2, j);
            const double h4 = 4.0 * q2(c - e);
            BT(j, 0, 0) = h4 * (10.0 * q2(e - c) + (e - c) * (f - e) + q2(f - e)) / (q2(c - g) * q2(e - g));
            BT(j, 0, 1) = h4 * (19.0 * q2(e - c) - (e - c) * (g - e) + 2.0 * (f - c) * ((f - c) + (g - e))) /
                          ((c - f) * q2(c - g) * (g - e));
            BT(j, 0, 2) = h4 * (10.0 * q2(e - c) + (e - c) * ((f - c) + (g - e)) + q2((f - c) + (g - e))) / (q2(c - f) * q2(c - g));
            BT(j, 1, 0) = h4 * (10.0 * q2(e - c) + q2(c - bb) + (c - bb) * (e - c)) / (q2(bb - f) * q2(c - f));
            BT(j, 1, 1) = h4 * ((c - e) * ((c - bb) + 20.0 * (e - c)) + (2.0 * (c - bb) + (e - c)) * (f - c)) /
                          ((e - bb) * q2(bb - f) * (f - c));
            BT(j, 1, 2) = h4 * (10.0 * q2(e - c) + (e - c) * (f - e) + q2(f - e)) / (q2(bb - e) * q2(bb - f));
            BT(j, 2, 0) = h4 * (12.0 * q2(e - c) + q2((c - a) + (c - bb)) + 3.0 * ((c - a) + (c - bb)) * (e - c)) /
                          (q2(a - e) * q2(bb - e));
            BT(j, 2, 1) = h4 * (19.0 * q2(e - c) + ((c - a) * (c - e)) + 2.0 * (e - bb) * ((c - a) + (e - bb))) /
                          ((a - c) * q2(a - e) * (e - bb));
            BT(j, 2, 2) = h4 * (10.0 * q2(e - c) + q2(c - bb) + (c - bb) * (e - c)) / (q2(a - c) * q2(a - e));
        }
        if (h->P.null_weights) {  // m_weno.fpp:345-361
            auto renorm = [&](bool right, int j) {
                double t = 0;
                for (int q = 0; q < 3; q++) t += right ? DR(q, j) : DL(q, j);
                for (int q = 0; q < 3; q++) { if (right) DR(q, j) /= t; else DL(q, j) /= t; }
            };
            if (bcb == MFC5_BC_RIEMANN_EXTRAP) {
                DR(1, 0) = 0; DR(2, 0) = 0; DR(0, 0) = 1; DL(1, 0) = 0; DL(2, 0) = 0; DL(0, 0) = 1;
                DR(2, 1) = 0; renorm(true, 1);
                DL(2, 1) = 0; renorm(false, 1);
            }
            if (bce == MFC5_BC_RIEMANN_EXTRAP) {
                DR(0, s - 1) = 0; renorm(true, s - 1);
                DL(0, s - 1) = 0; renorm(false, s - 1);
                DR(0, s) = 0; DR(1, s) = 0; DR(2, s) = 1; DL(0, s) = 0; DL(1, s) = 0; DL(2, s) = 1;
            }
        }
    }
    if (h->P.weno_order == 7) {
        for (int j = -b + pn; j <= s + b - pn; j++) {
            double xb[8], cl[wgen::KM][wgen::KM - 1], cr[wgen::KM][wgen::KM - 1], dl[wgen::KM], dr[wgen::KM], bt[wgen::KM][6];
            for (int m = 0; m < 8; m++) xb[m] = x[j - 4 + m];
            wgen::tables(4, xb, cl, cr, dl, dr, bt);
            for (int q = 0; q < 4; q++) {
                for (int r = 0; r < 3; r++) { CL(j, q, r) = cl[q][r]; CR(j, q, r) = cr[q][r]; }
                // the reference numbers the ideal weights of order 7 from the LEFTMOST stencil (d_cb(0) = 1/35 on a uniform grid)
                // while poly_coef / beta_coef keep the rightmost-first index: d(q) belongs to stencil 3 - q
                DL(q, j) = dl[3 - q];
                DR(q, j) = dr[3 - q];
                for (int r = 0; r < 6; r++) tab[4][((size_t)(j + b) * ns + q) * 6 + r] = bt[q][r];
            }
        }
    }
    uniform = 1;  // m_weno.fpp:869-879
    if (s > 0) {
        const double h0 = (x[s] - x[0]) / (double)s;
        for (int i = 0; i <= s - 1; i++)
            if (std::fabs((x[i + 1] - x[i]) - h0) > std::sqrt(2.220446049250313e-16) * std::fabs(h0)) { uniform = 0; break; }
    }
}

// ---- CBC coefficients: s_compute_cbc_coefficients m_cbc.fpp:327-408 ----------------------------------
static void cbc_tables(mfc5_handle* h, int d, int side) {
    const int s = h->s[d], b = h->b;
    const double* w = h->ds[d].data() + b;
    CbcCoef& K = h->cbc[d][side];
    std::memset(&K, 0, sizeof(K));
    double ds[8], x[10];
    for (int i = 0; i <= b; i++) ds[i] = side == 0 ? w[i] : w[s - i];
    x[0] = 0.0;
    for (int i = 0; i <= b; i++) x[i + 1] = x[i] + ds[i];
    if (h->P.weno_order == 3) {
        K.fd[0] = -6.0 / (3.0 * ds[0] + 2.0 * ds[1] - ds[2]);
        K.fd[1] = -4.0 * K.fd[0] / 3.0;
        K.fd[2] = K.fd[0] / 3.0;
        K.pi[0][0] = (x[0] - x[1]) / (x[0] - x[2]);
    } else {
        K.fd[0] = -50.0 / (25.0 * ds[0] + 2.0 * ds[1] - 1.e1 * ds[2] + 1.e1 * ds[3] - 3.0 * ds[4]);
        K.fd[1] = -48.0 * K.fd[0] / 25.0;
        K.fd[2] = 36.0 * K.fd[0] / 25.0;
        K.fd[3] = -16.0 * K.fd[0] / 25.0;
        K.fd[4] = 3.0 * K.fd[0] / 25.0;
        K.pi[0][0] = ((x[0] - x[1]) * (x[1] - x[2]) * (x[1] - x[3])) / ((x[1] - x[4]) * (x[4] - x[0]) * (x[4] - x[2]));
        K.pi[0][1] = ((x[1] - x[0]) * (x[1] - x[2]) * ((x[1] - x[3]) * (x[1] - x[3]) - (x[0] - x[4]) * ((x[3] - x[1]) + (x[4] - x[1])))) /
                     ((x[0] - x[3]) * (x[1] - x[3]) * (x[0] - x[4]) * (x[1] - x[4]));
        K.pi[0][2] = (x[1] - x[0]) * ((x[1] - x[2]) * (x[1] - x[3]) + ((x[0] - x[2]) + (x[1] - x[3])) * (x[0] - x[4])) /
                     ((x[2] - x[0]) * (x[0] - x[3]) * (x[0] - x[4]));
        K.pi[1][0] = ((x[0] - x[2]) * (x[2] - x[1]) * (x[2] - x[3])) / ((x[2] - x[4]) * (x[4] - x[0]) * (x[4] - x[1]));
        K.pi[1][1] = ((x[0] - x[2]) * (x[1] - x[2]) * ((x[1] - x[3]) * (x[2] - x[3]) + (x[0] - x[4]) * ((x[1] - x[3]) + (x[2] - x[4])))) /
                     ((x[0] - x[3]) * (x[1] - x[3]) * (x[0] - x[4]) * (x[1] - x[4]));
        K.pi[1][2] = ((x[1] - x[2]) * (x[2] - x[3]) * (x[2] - x[4])) / ((x[0] - x[2]) * (x[0] - x[3]) * (x[0] - x[4]));
    }
    K.ds0 = ds[0];
    double mx = w[-b];  // Del = maxval(d<xyz>) over the ghosted array, m_cbc.fpp:312-313
    for (int i = -b; i <= s + b; i++) mx = std::fmax(mx, w[i]);
    K.Del = mx;
    for (int a = 0; a < 3; a++) { K.vel_in[a] = h->P.vel_in[d][a]; K.vel_out[a] = h->P.vel_out[d][a]; }
    K.pres_in = h->P.pres_in[d]; K.pres_out = h->P.pres_out[d];
    for (int f = 0; f < MFC5_MAX_FLUIDS; f++) { K.arho_in[f] = h->P.alpha_rho_in[d][f]; K.alf_in[f] = h->P.alpha_in[d][f]; }
    K.grcbc_in = h->P.grcbc_in[d]; K.grcbc_out = h->P.grcbc_out[d]; K.grcbc_vel_out = h->P.grcbc_vel_out[d];
}

static int upload(double** dptr, const std::vector<double>& v) {
    if (*dptr) cudaFree(*dptr);
    *dptr = nullptr;
    CK(cudaMalloc(dptr, std::max<size_t>(v.size(), 1) * sizeof(double)));
    if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int mfc5_finalize_setup(mfc5_handle* h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->P.device));
    for (int d = 0; d < h->nd; d++)
        if (!h->grid_set[d]) return fail("mfc5_finalize_setup: grid not set for every direction");
    DevP& D = h->D;
    std::memset(&D, 0, sizeof(D));
    for (int d = 0; d < 3; d++) D.s[d] = h->s[d];
    D.nd = h->nd; D.nf = h->nf; D.N = h->N; D.b = h->b; D.pn = h->pn; D.ns = h->ns;
    D.wo = h->P.weno_order; D.wv = h->P.mapped_weno ? 1 : (h->P.wenoz ? 2 : 0); D.rs = h->P.riemann_solver;
    D.hll_u = h->P.hll_u_interface; D.avg = h->P.avg_state; D.ws = h->P.wave_speeds; D.mpp = h->P.mpp_lim;
    D.mixerr = h->P.mixture_err;
    // division-light arithmetic only where its assumptions hold: direct wave speeds, no weight nulling,
    // weno_eps large enough that products of four smoothness indicators cannot underflow
    D.low_mach = h->P.low_Mach;
    D.alt_ss = h->P.alt_soundspeed;
    D.teno = h->P.teno; D.teno_ct = h->P.teno_CT; D.mp_weno = h->P.mp_weno; D.wenoz_q = h->P.wenoz_q;
    D.int_comp = h->P.int_comp; D.ic_eps = h->P.ic_eps; D.ic_beta = h->P.ic_beta;
    D.ic_cosh = std::cosh(h->P.ic_beta); D.ic_tanh = std::tanh(h->P.ic_beta);
    D.muscl = h->P.muscl_order; D.muscl_lim = h->P.muscl_lim; D.muscl_eps = h->P.muscl_eps;
    // fast: division-light arithmetic (WENO, Riemann solve, flux difference); fast_riem: the subset the case-optimised kernels are
    // compiled for.  wave_speeds = 2 and low_Mach are Riemann-only features: division-light too, in the generic kernels
    D.fast = h->P.fast_math && !h->P.null_weights && h->P.weno_eps >= 1e-30 &&
             !h->P.alt_soundspeed && h->P.riemann_solver != 5 && h->P.muscl_order == 0 && !h->P.teno && !h->P.mp_weno && h->P.int_comp == 0 &&
             h->P.weno_order != 7 && !h->P.viscous;
    if (getenv("MFC5_NO_FAST")) D.fast = 0;
    D.fast_riem = D.fast && h->P.wave_speeds == 1 && h->P.low_Mach == 0;
    if (getenv("MFC5_NO_FAST_EXT") && !D.fast_riem) D.fast = 0;   // (A/B knob: the round-1 behaviour for these cases)
    D.viscous = h->P.viscous;
    for (int i = 0; i < 2; i++) {   // Re_size / Re_idx / Res_gs, m_global_parameters.fpp:781-812
        D.re_size[i] = 0;
        for (int f = 0; f < h->nf && h->P.viscous; f++)
            if (h->P.fluid_Re[f][i] > 0) { D.re_idx[i][D.re_size[i]] = f; D.res[i][D.re_size[i]] = h->P.fluid_Re[f][i]; D.re_size[i]++; }
    }
    D.sy = h->px; D.sz = h->px * h->ey; D.fs = h->fs;
    D.base = h->ox + h->oy * D.sy + h->oz * D.sz;
    D.eps = h->P.weno_eps;
    for (int f = 0; f < MFC5_MAX_FLUIDS; f++) { D.gam[f] = h->P.fluid_gamma[f]; D.pinf[f] = h->P.fluid_pi_inf[f]; D.qv[f] = h->P.fluid_qv[f]; }
    const bool hll1 = (D.rs == 1 && !D.hll_u);
    D.nface = h->P.weno_order == 5 ? 2 : 1;
    D.ncbc_comp = h->N + (hll1 ? h->nf : 1);
    for (int d = 0; d < h->nd; d++) {
        std::vector<double> tab[5];
        int uni = 1;
        weno_tables(h, d, tab, uni);
        h->uniform[d] = uni;
        for (int k = 0; k < 5; k++) if (upload(&h->d_tab[d][k], tab[k])) return -1;
        if (upload(&h->d_tab[d][5], h->ds[d])) return -1;
        if (upload(&h->d_tab[d][9], h->scc[d])) return -1;
        D.T[d].cc = h->d_tab[d][9];
        D.T[d].cL = h->d_tab[d][0]; D.T[d].cR = h->d_tab[d][1]; D.T[d].dL = h->d_tab[d][2]; D.T[d].dR = h->d_tab[d][3];
        D.T[d].bt = h->d_tab[d][4]; D.T[d].ds = h->d_tab[d][5]; D.T[d].uniform = uni;
        {   // combined rows for the division-light kernels (layout: DirTables::tab)
            const int nslot = h->s[d] + 1 + 2 * h->b, ns = h->ns, pr = h->pn > 0 ? h->pn : 1;
            std::vector<double> rows((size_t)nslot * MFC5_TAB, 0.0);
            for (int sl = 0; sl < nslot; sl++) {
                double* r = rows.data() + (size_t)sl * MFC5_TAB;
                for (int q = 0; q < ns && h->pn < 3; q++) {   // (order 7 runs the reference-order kernels: its rows only carry ds)
                    for (int k = 0; k < pr && h->pn > 0; k++) {
                        r[0 + q * pr + k] = tab[0][((size_t)sl * ns + q) * pr + k];
                        r[6 + q * pr + k] = tab[1][((size_t)sl * ns + q) * pr + k];
                    }
                    const double dl = tab[2][(size_t)sl * ns + q], dr = tab[3][(size_t)sl * ns + q];
                    r[12 + q] = dl; r[15 + q] = dr;
                    r[27 + q] = dl + dl * dl; r[30 + q] = dr + dr * dr;
                    r[33 + q] = 1.0 - 2.0 * dl; r[36 + q] = 1.0 - 2.0 * dr;
                    r[39 + q] = dl * dl; r[42 + q] = dr * dr;
                    for (int k = 0; k < 3; k++) r[18 + q * 3 + k] = tab[4][((size_t)sl * ns + q) * 3 + k];
                }
                r[45] = h->ds[d][sl];
                r[46] = 1.0 / h->ds[d][sl];
            }
            if (upload(&h->d_tab[d][6], rows)) return -1;
            D.T[d].tab = h->d_tab[d][6];
            std::vector<double> rowsT((size_t)nslot * MFC5_TAB);
            for (int sl = 0; sl < nslot; sl++)
                for (int k = 0; k < MFC5_TAB; k++) rowsT[(size_t)k * nslot + sl] = rows[(size_t)sl * MFC5_TAB + k];
            if (upload(&h->d_tab[d][7], rowsT)) return -1;
            D.T[d].tabT = h->d_tab[d][7];
            D.T[d].nslot = nslot;
            if (d == 0) {  // x kernel: [segment][MFC5_TAB / 2][lane][2], lane l of segment g = cell 30 g - 1 + l (clamped to s + 1)
                const int nseg = (h->s[0] + 1 + MFC5_XW - 1) / MFC5_XW;
                std::vector<double> rowsX((size_t)nseg * MFC5_TAB * 32);
                for (int g = 0; g < nseg; g++)
                    for (int l = 0; l < 32; l++) {
                        const int sl = std::min(MFC5_XW * g - 1 + l, h->s[0] + 1) + h->b;
                        for (int k = 0; k < MFC5_TAB; k++) rowsX[(size_t)g * MFC5_TAB * 32 + (size_t)(k >> 1) * 64 + 2 * l + (k & 1)] = rows[(size_t)sl * MFC5_TAB + k];
                    }
                if (upload(&h->d_tab[d][8], rowsX)) return -1;
                D.T[d].tabX = h->d_tab[d][8];
            }
        }
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        h->plane[d] = (long long)(h->s[t1] + 1) * (h->s[t2] + 1);
        for (int sd = 0; sd < 2; sd++) {
            D.bc[d][sd] = h->P.bc[d][sd];
            D.cbc_on[d][sd] = is_cbc(h->P.bc[d][sd]);
            if (D.cbc_on[d][sd]) {
                cbc_tables(h, d, sd);
                if (h->d_cbcbuf[d][sd]) cudaFree(h->d_cbcbuf[d][sd]);
                CK(cudaMalloc(&h->d_cbcbuf[d][sd], (size_t)(D.nface + 1) * D.ncbc_comp * h->plane[d] * sizeof(double)));
            }
            D.cbcbuf[d][sd] = h->d_cbcbuf[d][sd];
            D.dirich[d][sd] = h->d_dirich[d][sd];
        }
        for (int a = 0; a < 3; a++) { D.vb[d][a] = h->P.bc_vb[d][a]; D.ve[d][a] = h->P.bc_ve[d][a]; }
    }
    h->any_phys = h->any_nb = h->any_cbc = false;
    {
        CbcCoef all[6];
        std::memset(all, 0, sizeof(all));
        for (int d = 0; d < h->nd; d++)
            for (int sd = 0; sd < 2; sd++) {
                if (h->P.bc[d][sd] < 0) h->any_phys = true; else h->any_nb = true;
                if (D.cbc_on[d][sd]) { h->any_cbc = true; all[2 * d + sd] = h->cbc[d][sd]; }
            }
        CK(cudaMemcpy(h->d_cbccoef, all, sizeof(all), cudaMemcpyHostToDevice));
    }
    D.chunk = 0;
    if (const char* e = getenv("MFC5_CHUNK")) D.chunk = std::max(0, atoi(e));
    D.x_march = getenv("MFC5_X_MARCH") ? 1 : 0;
    h->use_graph = getenv("MFC5_NO_GRAPH") ? 0 : 1;
    h->use_side_stream = getenv("MFC5_NO_SIDE_STREAM") ? 0 : 1;
    for (int d = 0; d < h->nd; d++)
        if (h->P.bf[d] && h->P.bf_k[d] != 0.0) h->use_graph = 0;   // time-dependent acceleration: kernel arguments change every step
    h->warmed = false;
    h->setup = true;
    for (int i = 0; i < 4; i++) if (h->gexec[i]) { cudaGraphExecDestroy(h->gexec[i]); h->gexec[i] = nullptr; }
    h->prim_valid = false;
    h->pp = 0;
    // Case-optimised path with the second-generation marching kernels: extrapolation / reflective / periodic ghost cells are
    // read by the sweep loaders straight from the interior (no ghost-fill kernel for those faces), and the last sweep of a
    // stage writes the next stage's primitive variables over the rhs buffer (no cons -> prim pass, no extra buffer).
    const bool m2 = opt_march2(D);
    std::memset(D.gmode, 0, sizeof(D.gmode));
    if (m2 && !getenv("MFC5_NO_FOLD"))
        for (int d = 0; d < h->nd; d++)
            for (int sd = 0; sd < 2; sd++) {
                const int bc = h->P.bc[d][sd];
                if (bc == MFC5_BC_PERIODIC) D.gmode[d][sd] = MFC5_GM_WRAP;
                else if (bc == MFC5_BC_REFLECTIVE) D.gmode[d][sd] = MFC5_GM_MIRROR;
                else if (bc == MFC5_BC_GHOST_EXTRAP || is_cbc(bc)) D.gmode[d][sd] = MFC5_GM_CLAMP;   // ghost extrapolation; the CBC faces rewrite their fluxes
            }
    const bool was_alias = h->alias;
    h->alias = m2 && !getenv("MFC5_NO_FUSE_C2P");
    if (was_alias && !h->alias) h->d_prim[1] = nullptr;
    h->fuse_c2p = (h->alias || getenv("MFC5_FUSE_C2P")) ? 1 : 0;   // generic kernels: opt-in, separate buffer (see store_next_prim)
    if (h->alias) {
        if (h->d_prim[1] && h->d_prim[1] != h->d_rhs) cudaFree(h->d_prim[1]);
        h->d_prim[1] = h->d_rhs;
    } else if (h->fuse_c2p && !h->d_prim[1]) {
        const size_t nb = (size_t)h->N * h->fs * sizeof(double);
        CK(cudaMalloc(&h->d_prim[1], nb));
        CK(cudaMemset(h->d_prim[1], 0, nb));
    }
    return 0;
}

extern "C" int mfc5_set_dirichlet(mfc5_handle* h, int d, int side, const double* values) {
    if (!h || d < 0 || d >= h->nd || side < 0 || side > 1 || !values) return fail("mfc5_set_dirichlet: bad argument");
    CK(cudaSetDevice(h->P.device));
    const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    const size_t n = (size_t)h->N * (h->s[t1] + 1) * (h->s[t2] + 1);
    CK(cudaStreamSynchronize(h->st));
    if (!h->d_dirich[d][side]) {
        // first call for this face: the kernel argument changes from nullptr to the new buffer, so steps captured
        // before it (they hold DevP by value) must be re-captured.  Later calls reuse the allocation (its size is fixed).
        CK(cudaMalloc(&h->d_dirich[d][side], n * sizeof(double)));
        for (int i = 0; i < 4; i++) if (h->gexec[i]) { cudaGraphExecDestroy(h->gexec[i]); h->gexec[i] = nullptr; }
    }
    CK(cudaMemcpy(h->d_dirich[d][side], values, n * sizeof(double), cudaMemcpyHostToDevice));
    h->D.dirich[d][side] = h->d_dirich[d][side];
    return 0;
}

static int fill_full(mfc5_handle* h, const DevP& D);
static unsigned grid_for(long long n, int block) { return (unsigned)std::min<long long>((n + block - 1) / block, 148LL * 16); }

// ---- host <-> device field copies ------------------------------------------------------------------
static int copy_fields(mfc5_handle* h, double* dev, const double* const* hostc, double* const* hostm, int ghosted, bool to_dev, bool ghosts_valid = false) {
    CK(cudaSetDevice(h->P.device));
    const int gx = ghosted ? h->b : 0, gy = (ghosted && h->nd > 1) ? h->b : 0, gz = (ghosted && h->nd > 2) ? h->b : 0;
    // host extents
    const size_t hx = h->s[0] + 1 + 2 * gx, hy = h->s[1] + 1 + 2 * gy, hz = h->s[2] + 1 + 2 * gz;
    // copy region: upload -> interior only (ghost conserved values are never valid in the reference either); download ->
    // everything the host array holds, like GPU_UPDATE(host=...) of the whole sf array: the ghost cells of q_prim_vf are part
    // of its contract (m_rhs.fpp:884-898), those of q_cons_vf / rhs_vf receive the device's own (unused, zero) ghost cells
    const int cgx = to_dev ? 0 : gx, cgy = to_dev ? 0 : gy, cgz = to_dev ? 0 : gz;
    const size_t wx = h->s[0] + 1 + 2 * cgx, wy = h->s[1] + 1 + 2 * cgy, wz = h->s[2] + 1 + 2 * cgz;
    // Staging image = the other ping-pong state buffer: between calls it only holds the stale previous state
    // (stage 1 of the next step rewrites its whole interior), and one field of the host shape fits in one
    // device field (px >= m + 1 + 2b).  One contiguous DMA per field (one in total when the caller's fields
    // are adjacent in memory) + one repack kernel at HBM rate.
    double* stage = h->d_q[h->cur ^ 1];
    if (dev == stage) return fail("copy_fields: source aliases the staging buffer");
    const size_t hfs = hx * hy * hz;
    if ((long long)hfs > h->fs) return fail("copy_fields: staging image does not fit");
    RepackGeom G;
    G.N = h->N; G.wx = (int)wx; G.wy = (int)wy; G.wz = (int)wz;
    G.dev_off = (h->ox - cgx) + (long long)(h->oy - cgy) * h->px + (long long)(h->oz - cgz) * h->px * h->ey;
    G.sy = h->px; G.sz = h->px * h->ey; G.fs = h->fs;
    G.st_off = (gx - cgx) + (long long)(gy - cgy) * hx + (long long)(gz - cgz) * hx * hy;
    G.hx = hx; G.hxy = hx * hy; G.hfs = hfs;
    G.gx = cgx; G.gy = cgy; G.gz = cgz; G.zero_ghosts = (!to_dev && !ghosts_valid) ? 1 : 0;   // device ghost cells of q / rhs hold nothing meaningful
    bool adjacent = true;
    for (int v = 0; v < h->N; v++) {
        const double* hp = to_dev ? hostc[v] : hostm[v];
        if (!hp) return fail("null field pointer");
        const double* h0 = to_dev ? hostc[0] : hostm[0];
        if (hp != h0 + (size_t)v * hfs) adjacent = false;
    }
    const unsigned grid = grid_for((long long)wx * wy * wz * h->N, 256);
    if (!to_dev) { k_repack<<<grid, 256, 0, h->st>>>(dev, stage, G, 0); h->launches++; }
    if (adjacent) {
        if (to_dev) CK(cudaMemcpyAsync(stage, hostc[0], hfs * h->N * sizeof(double), cudaMemcpyHostToDevice, h->st));
        else CK(cudaMemcpyAsync(hostm[0], stage, hfs * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    } else {
        for (int v = 0; v < h->N; v++) {
            if (to_dev) CK(cudaMemcpyAsync(stage + (size_t)v * hfs, hostc[v], hfs * sizeof(double), cudaMemcpyHostToDevice, h->st));
            else CK(cudaMemcpyAsync(hostm[v], stage + (size_t)v * hfs, hfs * sizeof(double), cudaMemcpyDeviceToHost, h->st));
        }
    }
    if (to_dev) { k_repack<<<grid, 256, 0, h->st>>>(dev, stage, G, 1); h->launches++; }
    CK(cudaGetLastError());
    // stream-ordered on the compute stream (truly asynchronous for pinned host arrays); the call
    // returns when the data has arrived, like the reference's GPU_UPDATE
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int mfc5_upload_cons(mfc5_handle* h, const double* const* q, int ghosted) {
    if (!h || !q) return fail("mfc5_upload_cons: null argument");
    h->prim_valid = false;
    return copy_fields(h, h->d_q[h->cur], q, nullptr, ghosted, true);
}
extern "C" int mfc5_download_cons(mfc5_handle* h, double* const* q, int ghosted) {
    if (!h || !q) return fail("mfc5_download_cons: null argument");
    return copy_fields(h, h->d_q[h->cur], nullptr, q, ghosted, false);
}
static int refresh_prim(mfc5_handle* h);
static long long halo_count(const mfc5_handle* h, int d, bool full);
extern "C" int mfc5_download_prim(mfc5_handle* h, double* const* q, int ghosted) {
    if (!h || !q) return fail("mfc5_download_prim: null argument");
    if (h->setup) {
        CK(cudaSetDevice(h->P.device));
        // after a time step the primitive buffer still holds the last STAGE's input: convert the current state first
        if (!h->prim_valid && refresh_prim(h)) return -1;
        if (ghosted) {  // edges and corners as the reference fills them (the sweeps do not need them)
            DevP D = h->D;
            D.prim = h->d_prim[h->pp]; D.primw = h->d_prim[h->pp];
            if (fill_full(h, D)) return -1;
        }
    }
    return copy_fields(h, h->d_prim[h->pp], nullptr, q, ghosted, false, true);
}
extern "C" int mfc5_download_rhs(mfc5_handle* h, double* const* q, int ghosted) {
    if (!h || !q) return fail("mfc5_download_rhs: null argument");
    return copy_fields(h, rhs_ptr(h), nullptr, q, ghosted, false);
}

// ---- stage orchestration ---------------------------------------------------------------------------
template <typename F>
static int by_shape(const mfc5_handle* h, F&& f) {
    switch (h->nf * 10 + h->nd) {
    case 11: f(std::integral_constant<int, 1>(), std::integral_constant<int, 1>()); return 0;
    case 12: f(std::integral_constant<int, 1>(), std::integral_constant<int, 2>()); return 0;
    case 13: f(std::integral_constant<int, 1>(), std::integral_constant<int, 3>()); return 0;
    case 21: f(std::integral_constant<int, 2>(), std::integral_constant<int, 1>()); return 0;
    case 22: f(std::integral_constant<int, 2>(), std::integral_constant<int, 2>()); return 0;
    case 23: f(std::integral_constant<int, 2>(), std::integral_constant<int, 3>()); return 0;
    case 31: f(std::integral_constant<int, 3>(), std::integral_constant<int, 1>()); return 0;
    case 32: f(std::integral_constant<int, 3>(), std::integral_constant<int, 2>()); return 0;
    case 33: f(std::integral_constant<int, 3>(), std::integral_constant<int, 3>()); return 0;
    }
    return fail("unsupported (num_fluids, num_dims)");
}


// timing segments: when h->timing, every kernel class is bracketed by a pair of events
static int seg_begin(mfc5_handle* h, int cls, cudaStream_t st = nullptr) {
    if (!h->timing) return 0;
    if (!st) st = h->st;
    const size_t i = h->seg_cls.size();
    while (h->ev_pool.size() < 2 * (i + 1)) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev_pool.push_back(e);
    }
    h->seg_cls.push_back(cls);
    CK(cudaEventRecord(h->ev_pool[2 * i], st));
    return 0;
}
static int seg_end(mfc5_handle* h, int nlaunch, cudaStream_t st = nullptr) {
    h->launches += nlaunch;
    if (!h->timing) return 0;
    if (!st) st = h->st;
    const size_t i = h->seg_cls.size() - 1;
    h->cls_launches[h->seg_cls[i]] += nlaunch;
    CK(cudaEventRecord(h->ev_pool[2 * i + 1], st));
    return 0;
}

// doubles in one packed halo face of direction d: b layers x transverse interior x N variables
// doubles per packed face of direction d; full: with the ghost layers of the directions before d (viscous path)
static long long halo_count(const mfc5_handle* h, int d, bool full) {
    long long n = (long long)h->N * h->b;
    for (int e = 0; e < 3; e++)
        if (e != d) n *= (h->s[e] + 1 + ((full && e < d) ? 2 * h->b : 0));
    return n;
}
// buffer size: the larger of the two layouts when the case can take the viscous path
static size_t halo_bytes(const mfc5_handle* h, int d) { return (size_t)halo_count(h, d, h->P.viscous != 0) * sizeof(double); }

static unsigned faces_grid_x(const mfc5_handle* h, long long per_thread_items) {
    long long mx = 1;
    for (int d = 0; d < h->nd; d++) {
        long long n = per_thread_items;
        for (int e = 0; e < 3; e++)
            if (e != d) n *= (h->s[e] + 1);
        mx = std::max(mx, n);
    }
    return grid_for(mx, 256);
}

// reference-order ghost fill (edges and corners included) of the physical faces, for callers that ask for
// the ghosted q_prim_vf (m_rhs.fpp:884-898); processor-face ghosts hold the transverse interior only
static int fill_full(mfc5_handle* h, const DevP& D) {
    for (int d = 0; d < h->nd; d++) {
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        long long n1 = h->s[t1] + 1, n2 = h->s[t2] + 1;
        if (t1 < d) n1 += 2 * h->b;
        if (t2 < d && t2 < h->nd) n2 += 2 * h->b;
        for (int sd = 0; sd < 2; sd++)
            if (h->P.bc[d][sd] < 0) {
                k_bc_fill<<<grid_for(n1 * n2 * h->N, 256), 256, 0, h->st>>>(D, d, sd, 1);
                h->launches++;
            }
    }
    CK(cudaGetLastError());
    return 0;
}

// halo exchange of every processor face on the communication stream: one pack launch, one NCCL group,
// one unpack launch (replaces s_mpi_sendrecv_variables_buffers, m_mpi_common.fpp:549-1070).  The faces of
// the three directions are independent because only the transverse interior travels.
static int exchange_halos(mfc5_handle* h, const DevP& D) {
    if (!h->comm) return fail("processor boundary (bc >= 0) but mfc5_comm_init was not called");
    cudaStream_t sc = h->st_comm;
    HaloBufs snd, rcv;
    for (int d = 0; d < 3; d++)
        for (int sd = 0; sd < 2; sd++) {
            snd.buf[2 * d + sd] = rcv.buf[2 * d + sd] = nullptr;
            if (d >= h->nd || h->P.bc[d][sd] < 0) continue;
            if (!h->d_send[d][sd]) {
                const size_t nb = halo_bytes(h, d);
                CK(cudaMalloc(&h->d_send[d][sd], nb));
                CK(cudaMalloc(&h->d_recv[d][sd], nb));
            }
            snd.buf[2 * d + sd] = h->d_send[d][sd];
            rcv.buf[2 * d + sd] = h->d_recv[d][sd];
        }
    const dim3 grid(faces_grid_x(h, (long long)h->N * h->b), 6, 1);
    CK(cudaStreamWaitEvent(sc, h->ev_prim, 0));
    if (seg_begin(h, MFC5_K_HALO, sc)) return -1;
    k_halo_faces<<<grid, 256, 0, sc>>>(D, snd, 1);
    NK(g_nccl.GroupStart());
    // per direction: send low face then high face, receive high then low, so that the pairing is also
    // right when both neighbours are the same rank (periodic direction split over 2 ranks)
    for (int d = 0; d < h->nd; d++) {
        const size_t cnt = (size_t)halo_count(h, d, false);
        for (int sd = 0; sd < 2; sd++)
            if (h->P.bc[d][sd] >= 0) NK(g_nccl.Send(h->d_send[d][sd], cnt, NCCL_FLOAT64, h->P.bc[d][sd], h->comm, sc));
        for (int sd = 1; sd >= 0; sd--)
            if (h->P.bc[d][sd] >= 0) NK(g_nccl.Recv(h->d_recv[d][sd], cnt, NCCL_FLOAT64, h->P.bc[d][sd], h->comm, sc));
    }
    NK(g_nccl.GroupEnd());
    k_halo_faces<<<grid, 256, 0, sc>>>(D, rcv, 0);
    if (seg_end(h, 2, sc)) return -1;
    CK(cudaEventRecord(h->ev_halo, sc));
    return 0;
}

// Viscous path: ghost cells INCLUDING edges and corners (the velocity gradients of a face read them), direction by direction
// on the compute stream as s_populate_variables_buffers does (m_boundary_common.fpp:99-187): the physical fills of direction d
// and its halo exchange both run over the ghost layers of the directions before d.
static int fill_exchange_full(mfc5_handle* h, const DevP& D, int* nlaunch) {
    int nl = 0;
    for (int d = 0; d < h->nd; d++) {
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        long long n1 = h->s[t1] + 1, n2 = h->s[t2] + 1;
        if (t1 < d) n1 += 2 * h->b;
        if (t2 < d && t2 < h->nd) n2 += 2 * h->b;
        for (int sd = 0; sd < 2; sd++)
            if (h->P.bc[d][sd] < 0) {
                k_bc_fill<<<grid_for(n1 * n2 * h->N, 256), 256, 0, h->st>>>(D, d, sd, 1);
                nl++;
            }
        if (h->P.bc[d][0] < 0 && h->P.bc[d][1] < 0) continue;
        if (!h->comm) return fail("processor boundary (bc >= 0) but mfc5_comm_init was not called");
        HaloBufs snd, rcv;
        for (int f = 0; f < 6; f++) snd.buf[f] = rcv.buf[f] = nullptr;
        for (int sd = 0; sd < 2; sd++) {
            if (h->P.bc[d][sd] < 0) continue;
            if (!h->d_send[d][sd]) {
                CK(cudaMalloc(&h->d_send[d][sd], halo_bytes(h, d)));
                CK(cudaMalloc(&h->d_recv[d][sd], halo_bytes(h, d)));
            }
            snd.buf[2 * d + sd] = h->d_send[d][sd];
            rcv.buf[2 * d + sd] = h->d_recv[d][sd];
        }
        const size_t cnt = (size_t)halo_count(h, d, true);
        const dim3 grid(grid_for((long long)cnt, 256), 2, 1);
        k_halo_faces<<<grid, 256, 0, h->st>>>(D, snd, 1, 1, 2 * d);
        NK(g_nccl.GroupStart());
        for (int sd = 0; sd < 2; sd++)
            if (h->P.bc[d][sd] >= 0) NK(g_nccl.Send(h->d_send[d][sd], cnt, NCCL_FLOAT64, h->P.bc[d][sd], h->comm, h->st));
        for (int sd = 1; sd >= 0; sd--)
            if (h->P.bc[d][sd] >= 0) NK(g_nccl.Recv(h->d_recv[d][sd], cnt, NCCL_FLOAT64, h->P.bc[d][sd], h->comm, h->st));
        NK(g_nccl.GroupEnd());
        k_halo_faces<<<grid, 256, 0, h->st>>>(D, rcv, 0, 1, 2 * d);
        nl += 2;
    }
    CK(cudaGetLastError());
    *nlaunch = nl;
    return 0;
}

static int launch_sweep_raw(mfc5_handle* h, const DevP& D, int dir) {
    bool ok = false;
    if (!getenv("MFC5_NO_OPT")) ok = launch_sweep_opt(D, dir, h->st);
    if (!ok) ok = launch_sweep_generic_nf1(D, dir, h->st) || launch_sweep_generic_nf2(D, dir, h->st) || launch_sweep_generic_nf3(D, dir, h->st);
    if (!ok) return fail("no sweep kernel for this configuration");
    cudaError_t e_ = cudaPeekAtLastError();
    if (e_ != cudaSuccess) {
        cudaGetLastError();
        return fail(std::string("sweep launch failed, direction ") + std::to_string(dir) + ": " + cudaGetErrorString(e_));
    }
    return 0;
}

// the sweep of one direction.  With processor boundaries the x sweep (first of the stage) is split: the
// segments that touch no ghost cell start at once and overlap the halo exchange running on the
// communication stream; the edge segments and the later sweeps wait for it.
static int launch_sweep(mfc5_handle* h, const DevP& D0, int dir, bool* halo_pending) {
    DevP D = D0;
    int nl = 0;
    if (seg_begin(h, MFC5_K_SWEEP_X + dir)) return -1;
    const bool x_shuffle = dir == 0 && !(D.x_march && (long long)(h->s[1] + 1) * (h->s[2] + 1) >= 4096 && h->nd >= 2);
    if (*halo_pending && x_shuffle) {
        const int nx = h->s[0] + 1, segs = (nx + MFC5_XW - 1) / MFC5_XW;
        // segment g reads cells [30 g - 1 - w, min(30 g + 30 + w, m + 1 + w)], w = 2 (3 at order 7): interior iff that range is within [0, m]
        const int wh = h->pn == 3 ? 3 : 2;
        int lo = 0, hi = segs;  // interior segments [lo, hi)
        while (lo < segs && MFC5_XW * lo - 1 - wh < 0) lo++;
        while (hi > lo && std::min(MFC5_XW * (hi - 1) + 30 + wh, h->s[0] + 1 + wh) > h->s[0]) hi--;
        D.xseg_hole = segs; D.xseg_hole_n = 0;
        if (hi > lo) {
            D.xseg_lo = lo; D.xseg_n = hi - lo;
            if (launch_sweep_raw(h, D, 0)) return -1;
            nl++;
        }
        CK(cudaStreamWaitEvent(h->st, h->ev_halo, 0));
        *halo_pending = false;
        if (lo > 0 || hi < segs) {   // both ends in ONE launch (one tail instead of two): the segment list minus the interior ones
            D.xseg_lo = 0; D.xseg_n = segs - (hi - lo);
            D.xseg_hole = lo; D.xseg_hole_n = hi - lo;
            if (launch_sweep_raw(h, D, 0)) return -1;
            nl++;
        }
    } else {
        if (*halo_pending) {
            CK(cudaStreamWaitEvent(h->st, h->ev_halo, 0));
            *halo_pending = false;
        }
        D.xseg_lo = 0; D.xseg_n = 0; D.xseg_hole = 1 << 30; D.xseg_hole_n = 0;
        if (launch_sweep_raw(h, D, dir)) return -1;
        nl++;
    }
    return seg_end(h, nl);
}

// one RHS evaluation (+ fused RK update when do_rk): q1 = state read, q2 = stage-1 copy, qout = target
static int run_stage(mfc5_handle* h, const double* q1, const double* q2, double* qout, int do_rk, const double rk[4], double dt) {
    NvtxRange nv_rhs("COMPUTE-RHS");
    DevP D = h->D;
    D.prim = h->d_prim[h->pp]; D.primw = h->d_prim[h->pp]; D.q1 = q1; D.q2 = q2; D.qout = qout; D.rhs = rhs_ptr(h);
    D.prim_next = (do_rk && h->fuse_c2p) ? h->d_prim[h->pp ^ 1] : nullptr;   // alias mode: == D.rhs
    D.bf_on = 0;
    if (do_rk && (h->P.bf[0] || h->P.bf[1] || h->P.bf[2])) {   // s_compute_acceleration(mytime), m_body_forces.fpp:206-218
        D.bf_on = 1;
        for (int d = 0; d < 3; d++)
            D.bf_a[d] = (d < h->nd && h->P.bf[d]) ? h->P.bf_g[d] + h->P.bf_k[d] * std::sin(h->P.bf_w[d] * h->mytime - h->P.bf_p[d]) : 0.0;
        D.bf_ldt = rk[2] * dt / rk[3];
    }
    D.rk0 = rk[0]; D.rk1 = rk[1]; D.rk2 = rk[2]; D.rk3 = rk[3]; D.irk3 = 1.0 / rk[3]; D.dt = dt; D.do_rk = do_rk;
    const long long ncell = (long long)(h->s[0] + 1) * (h->s[1] + 1) * (h->s[2] + 1);
    if (!h->prim_valid) {  // with MFC5_FUSE_C2P only after an upload: the previous stage wrote them
        NvtxRange nv("RHS-CONVERT");
        if (seg_begin(h, MFC5_K_CONS2PRIM)) return -1;
        int rc = by_shape(h, [&](auto nf, auto nd) {
            k_cons2prim<decltype(nf)::value, decltype(nd)::value><<<grid_for(ncell, 256), 256, 0, h->st>>>(D);
        });
        if (rc) return rc;
        if (seg_end(h, 1)) return -1;
    }
    bool halo_pending = false;
    NvtxRange nv_comm("RHS-COMMUNICATION");
    if (h->any_nb && !h->P.viscous) {
        CK(cudaEventRecord(h->ev_prim, h->st));
        if (exchange_halos(h, D)) return -1;
        halo_pending = true;
    }
    // Ghost fill and characteristic-BC fluxes, one launch per kind and face group.  Only the x faces are needed by the
    // x sweep: the y / z faces (written outside everything the x sweep reads) run on a side stream concurrently with it
    // and are joined before the y sweep.
    bool any_phys_x = false, any_phys_yz = false, any_cbc_x = false, any_cbc_yz = false;
    for (int d = 0; d < h->nd; d++)
        for (int sd = 0; sd < 2; sd++) {
            if (h->P.bc[d][sd] < 0 && D.gmode[d][sd] == MFC5_GM_MEM) (d == 0 ? any_phys_x : any_phys_yz) = true;
            if (D.cbc_on[d][sd]) (d == 0 ? any_cbc_x : any_cbc_yz) = true;
        }
    long long mx = 1;
    for (int d = 0; d < h->nd; d++) mx = std::max(mx, h->plane[d]);
    const CbcCoef* K = h->d_cbccoef;
    auto faces = [&](int face0, int nfaces, bool phys, bool cbc, cudaStream_t st) -> int {
        if (phys) {
            if (seg_begin(h, MFC5_K_BC_FILL, st)) return -1;
            k_bc_faces<<<dim3(faces_grid_x(h, h->N), nfaces, 1), 256, 0, st>>>(D, face0);
            if (seg_end(h, 1, st)) return -1;
        }
        if (cbc) {
            if (seg_begin(h, MFC5_K_CBC, st)) return -1;
            by_shape(h, [&](auto nf, auto nd) {
                k_cbc<decltype(nf)::value, decltype(nd)::value><<<dim3((unsigned)((mx + 127) / 128), nfaces, 1), 128, 0, st>>>(D, K, face0);
            });
            if (seg_end(h, 1, st)) return -1;
        }
        return 0;
    };
    bool side_pending = false;
    if (h->P.viscous) {
        // the velocity gradients of the viscous fluxes read edge / corner ghost cells: fill in the reference's order (x, then y
        // over the x ghosts, then z over both), m_boundary_common.fpp:143-149; the CBC fluxes follow on the same stream
        if (h->any_phys || h->any_nb) {
            if (seg_begin(h, MFC5_K_BC_FILL)) return -1;
            int nfill = 0;
            if (fill_exchange_full(h, D, &nfill)) return -1;
            if (seg_end(h, nfill)) return -1;
        }
        if (faces(0, 2 * h->nd, false, any_cbc_x || any_cbc_yz, h->st)) return -1;
        any_phys_x = any_phys_yz = any_cbc_x = any_cbc_yz = false;
    }
    if (h->nd > 1 && (any_phys_yz || any_cbc_yz)) {
        if (h->use_side_stream) {
            if (!h->any_nb) CK(cudaEventRecord(h->ev_prim, h->st));
            CK(cudaStreamWaitEvent(h->st_bc, h->ev_prim, 0));
            if (faces(2, 2 * (h->nd - 1), any_phys_yz, any_cbc_yz, h->st_bc)) return -1;
            CK(cudaEventRecord(h->ev_bc, h->st_bc));
            side_pending = true;
        } else if (faces(2, 2 * (h->nd - 1), any_phys_yz, any_cbc_yz, h->st)) return -1;
    }
    if (faces(0, 2, any_phys_x, any_cbc_x, h->st)) return -1;
    nv_comm.end();
    static const char* sweep_names[3] = {"RHS-RECONSTRUCTION+RIEMANN-SOLVER+ADVECTION-SRC x", "RHS-RECONSTRUCTION+RIEMANN-SOLVER+ADVECTION-SRC y",
                                         "RHS-RECONSTRUCTION+RIEMANN-SOLVER+ADVECTION-SRC z"};
    for (int d = 0; d < h->nd; d++) {
        NvtxRange nv(sweep_names[d]);
        if (d == 1 && side_pending) CK(cudaStreamWaitEvent(h->st, h->ev_bc, 0));
        if (launch_sweep(h, D, d, &halo_pending)) return -1;
    }
    // the next stage's cons2prim overwrites prim: the communication stream must be done reading it (it is:
    // every sweep waited for ev_halo), and the next pack must not start before this stage's sweeps finished
    // reading the ghosts it will overwrite -- ev_prim of the next stage is recorded after them on h->st.
    CK(cudaGetLastError());
    if (do_rk) {
        if (D.prim_next) { h->pp ^= 1; h->prim_valid = true; }   // primitives of the new state are in the other buffer
        else h->prim_valid = false;
    } else {
        h->prim_valid = true;                                     // d_prim[pp] holds the primitives of the (unchanged) state
    }
    return 0;
}

// primitive variables of the CURRENT conservative state into d_prim[pp] (interior; physical ghost faces are filled by the
// callers that need them)
static int refresh_prim(mfc5_handle* h) {
    DevP D = h->D;
    D.prim = h->d_prim[h->pp]; D.primw = h->d_prim[h->pp]; D.q1 = h->d_q[h->cur]; D.q2 = D.q1; D.qout = nullptr; D.rhs = rhs_ptr(h);
    const long long ncell = (long long)(h->s[0] + 1) * (h->s[1] + 1) * (h->s[2] + 1);
    int rc = by_shape(h, [&](auto nf, auto nd) {
        k_cons2prim<decltype(nf)::value, decltype(nd)::value><<<grid_for(ncell, 256), 256, 0, h->st>>>(D);
    });
    if (rc) return rc;
    h->launches += 1;
    CK(cudaGetLastError());
    h->prim_valid = true;
    return 0;
}

static const double RK3[3][4] = {{1, 0, 1, 1}, {1, 3, 1, 4}, {2, 1, 2, 3}};  // m_time_steppers.fpp:421-432
static const double RK2[2][4] = {{1, 0, 1, 1}, {1, 1, 1, 2}};
static const double RK1[1][4] = {{1, 0, 1, 1}};

// enqueue one full step on h->st.  Ping-pong: stage 1 reads A (which thereby *is* q_cons_ts(2)) and
// writes B; later stages update B in place.  For time_stepper = 1 the update is in place.
static int enqueue_step(mfc5_handle* h, double dt, bool timed) {
    NvtxRange nv_step("TIMESTEP");
    const int ns = h->P.time_stepper;
    double* A = h->d_q[h->cur];
    double* B = ns > 1 ? h->d_q[h->cur ^ 1] : A;
    h->timing = timed;
    if (timed) {
        h->seg_cls.clear();
        for (int k = 0; k < MFC5_NUM_KCLASS; k++) { h->cls_ms[k] = 0.0; h->cls_launches[k] = 0; }
    }
    for (int st = 1; st <= ns; st++) {
        const double* rk = ns == 3 ? RK3[st - 1] : ns == 2 ? RK2[st - 1] : RK1[0];
        const double* q1 = st == 1 ? A : B;
        if (timed) CK(cudaEventRecord(h->ev_stage[st - 1], h->st));
        if (run_stage(h, q1, A, B, 1, rk, dt)) { h->timing = false; return -1; }
    }
    if (timed) CK(cudaEventRecord(h->ev_stage[ns], h->st));
    h->timing = false;
    if (ns > 1) h->cur ^= 1;
    h->mytime = h->mytime + dt;   // m_start_up.fpp:634
    return 0;
}

extern "C" int mfc5_compute_rhs(mfc5_handle* h) {
    if (!h || !h->setup) return fail("mfc5_compute_rhs: handle not set up");
    CK(cudaSetDevice(h->P.device));
    const double rk[4] = {1, 0, 1, 1};
    if (run_stage(h, h->d_q[h->cur], h->d_q[h->cur], h->d_q[h->cur], 0, rk, 0.0)) return -1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int mfc5_tvd_rk(mfc5_handle* h, double dt, double* stage_time_min) {
    if (!h || !h->setup) return fail("mfc5_tvd_rk: handle not set up");
    CK(cudaSetDevice(h->P.device));
    if (enqueue_step(h, dt, true)) return -1;
    h->warmed = true;
    CK(cudaStreamSynchronize(h->st));
    const int ns = h->P.time_stepper;
    for (int st = 0; st < ns; st++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev_stage[st], h->ev_stage[st + 1]));
        if (stage_time_min) {
            const double t = ms * 1e-3;
            *stage_time_min = (*stage_time_min <= 0.0) ? t : std::min(*stage_time_min, t);
        }
    }
    for (size_t i = 0; i < h->seg_cls.size(); i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
        h->cls_ms[h->seg_cls[i]] += ms;
    }
    return 0;
}

// Enqueue `nsteps` steps without host synchronisation.  After one uncaptured step the whole step (every
// kernel of the 3 stages, the communication-stream fork/join and the NCCL group) is captured once per
// ping-pong parity into a CUDA graph and replayed: a step costs one graph launch instead of ~15-50 kernel
// launches, which is what matters on the small configurations (C1, C2, C4 strong scaling).
extern "C" int mfc5_tvd_rk_n(mfc5_handle* h, double dt, int nsteps) {
    if (!h || !h->setup) return fail("mfc5_tvd_rk_n: handle not set up");
    CK(cudaSetDevice(h->P.device));
    for (int i = 0; i < nsteps; i++) {
        // eager: the first step (lazy allocations, function attributes) and any step that has to start with a
        // cons -> prim pass (right after an upload), which the captured sequence does not contain
        if (!h->use_graph || !h->warmed || (h->fuse_c2p && !h->prim_valid)) {
            if (enqueue_step(h, dt, false)) return -1;
            h->warmed = true;
            continue;
        }
        // Outside MFC5_FUSE_C2P mode a captured step must always start with its cons -> prim pass: whether stage 1 launches
        // it depends on prim_valid at CAPTURE time (mfc5_reduce / mfc5_compute_rhs leave it true), and a graph captured
        // without it would replay on stale primitive variables.  The pass is redundant right after those calls, but correct.
        if (!h->fuse_c2p) h->prim_valid = false;
        const int cur0 = h->cur, pp0 = h->pp, key = 2 * cur0 + pp0;  // buffer parities the captured pointers depend on
        if (!h->gexec[key] || h->gdt[key] != dt) {
            if (h->gexec[key]) { cudaGraphExecDestroy(h->gexec[key]); h->gexec[key] = nullptr; }
            const long long l0 = h->launches;
            const bool pv0 = h->prim_valid;
            const double t0 = h->mytime;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(h->st, cudaStreamCaptureModeRelaxed);
            int rc = 0;
            if (e == cudaSuccess) {
                rc = enqueue_step(h, dt, false);
                e = cudaStreamEndCapture(h->st, &graph);
            }
            const int cur1 = h->cur, pp1 = h->pp;  // where a step leaves the parities
            h->launches_captured = h->launches - l0;
            if (e == cudaSuccess && rc == 0 && graph) e = cudaGraphInstantiate(&h->gexec[key], graph, 0);
            if (graph) cudaGraphDestroy(graph);
            h->launches = l0;
            h->cur = cur0; h->pp = pp0; h->prim_valid = pv0; h->mytime = t0;  // the captured enqueue only did the bookkeeping, nothing ran
            if (e != cudaSuccess || rc != 0 || !h->gexec[key]) {
                // capture not possible here (e.g. an NCCL build without graph support): run eagerly from now on
                g_err = std::string("mfc5_tvd_rk_n: CUDA graph capture failed (") + (e != cudaSuccess ? cudaGetErrorString(e) : "enqueue error") +
                        "); running eagerly from now on";   // not fatal: readable through mfc5_last_error / mfc5_graph_state
                cudaGetLastError();
                h->gexec[key] = nullptr;
                h->use_graph = 0;
                h->graph_fallback = true;
                if (enqueue_step(h, dt, false)) return -1;
                continue;
            }
            h->glaunches[key] = h->launches_captured;
            h->gdt[key] = dt;
            h->gcur1[key] = cur1; h->gpp1[key] = pp1;
        }
        CK(cudaGraphLaunch(h->gexec[key], h->st));
        h->launches += h->glaunches[key];
        h->cur = h->gcur1[key]; h->pp = h->gpp1[key];
        h->prim_valid = h->fuse_c2p != 0;
        h->mytime = h->mytime + dt;
    }
    return 0;
}

// 1: steps enqueued by mfc5_tvd_rk_n replay a captured CUDA graph; 0: eager by request (MFC5_NO_GRAPH, time-dependent body
// force); -1: a capture failed and the handle fell back to eager launches (the reason is in mfc5_last_error at that time)
extern "C" int mfc5_graph_state(const mfc5_handle* h) {
    if (!h) return 0;
    return h->graph_fallback ? -1 : (h->use_graph ? 1 : 0);
}

extern "C" int mfc5_set_time(mfc5_handle* h, double mytime) {
    if (!h) return fail("null handle");
    h->mytime = mytime;
    return 0;
}

extern "C" int mfc5_sync(mfc5_handle* h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->P.device));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int mfc5_reduce(mfc5_handle* h, int what, double dt, double* out) {
    if (!h || !h->setup || !out) return fail("mfc5_reduce: bad argument");
    if (what < 0 || what > 2) return fail("mfc5_reduce: what must be 0, 1 or 2");
    CK(cudaSetDevice(h->P.device));
    // primitive variables of the current state
    DevP D = h->D;
    D.prim = h->d_prim[h->pp]; D.primw = h->d_prim[h->pp]; D.q1 = h->d_q[h->cur]; D.q2 = D.q1; D.qout = nullptr; D.rhs = rhs_ptr(h);
    h->prim_valid = true;  // the cons -> prim pass below refreshes d_prim[pp] for the current state
    const long long ncell = (long long)(h->s[0] + 1) * (h->s[1] + 1) * (h->s[2] + 1);
    const unsigned long long init[4] = {0ull, 0x7FEFFFFFFFFFFFFFull, 0ull, 0ull};
    CK(cudaMemcpyAsync(h->d_red, init, sizeof(init), cudaMemcpyHostToDevice, h->st));
    int rc = by_shape(h, [&](auto nf, auto nd) {
        k_cons2prim<decltype(nf)::value, decltype(nd)::value><<<grid_for(ncell, 256), 256, 0, h->st>>>(D);
        k_reduce<decltype(nf)::value, decltype(nd)::value><<<grid_for(ncell, 256), 256, 0, h->st>>>(D, dt, h->d_red);
    });
    if (rc) return rc;
    h->launches += 2;
    unsigned long long res[4];
    if (h->comm) {
        // max / min / sum of three scalars in ONE collective: gather every rank's 4 words, combine on the host (the three
        // separate 1-element all-reduces cost three NCCL latencies per cfl_adap_dt step)
        NK(g_nccl.AllGather(h->d_red, h->d_red_all, 4, NCCL_UINT64, h->comm, h->st));
        std::vector<unsigned long long> all((size_t)4 * h->nranks);
        CK(cudaMemcpyAsync(all.data(), h->d_red_all, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        res[0] = 0ull; res[1] = 0x7FEFFFFFFFFFFFFFull; res[2] = 0ull; res[3] = 0ull;
        for (int r = 0; r < h->nranks; r++) {   // non-negative doubles order like their bit patterns
            res[0] = std::max(res[0], all[4 * r + 0]);
            res[1] = std::min(res[1], all[4 * r + 1]);
            res[2] += all[4 * r + 2];
        }
    } else {
        CK(cudaMemcpyAsync(res, h->d_red, sizeof(res), cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
    }
    if (what == 2) { *out = (double)res[2]; return 0; }
    double v;
    std::memcpy(&v, &res[what], sizeof(double));
    if (what == 0 && res[2] > 0) v = NAN;  // a NaN ICFL must not be hidden by fmax
    *out = v;
    return 0;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------
extern "C" int mfc5_nccl_unique_id(char id_out[128]) {
    if (nccl_load()) return -1;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, id.internal, 128);
    return 0;
}

extern "C" int mfc5_comm_init(mfc5_handle* h, const char id_in[128], int rank, int nranks) {
    if (!h) return fail("null handle");
    if (nccl_load()) return -1;
    CK(cudaSetDevice(h->P.device));
    ncclUniqueId id;
    std::memcpy(id.internal, id_in, 128);
    NK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks;
    if (h->d_red_all) cudaFree(h->d_red_all);
    CK(cudaMalloc(&h->d_red_all, (size_t)4 * nranks * sizeof(unsigned long long)));
    // packed halo faces: allocated here, not lazily inside the first stage
    for (int d = 0; d < h->nd; d++)
        for (int sd = 0; sd < 2; sd++)
            if (h->P.bc[d][sd] >= 0 && !h->d_send[d][sd]) {
                const size_t nb = halo_bytes(h, d);
                CK(cudaMalloc(&h->d_send[d][sd], nb));
                CK(cudaMalloc(&h->d_recv[d][sd], nb));
            }
    return 0;
}
