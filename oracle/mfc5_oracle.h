/* mfc5_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, not a product path).
 *
 * A plain-C, unfused, array-for-array restatement of the reference's 5-equation RHS + TVD-RK path
 * (MFlowCode/MFC @ 0b234b19).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Parity is PINNED: tests/test_oracle_golden.py runs
 * it against the reference's own golden files (tests/golden/<UUID>.npz) at the reference's
 * tolerance (toolchain/mfc/packer/tol.py:10-20).
 */
#ifndef MFC5_ORACLE_H
#define MFC5_ORACLE_H
#include "../include/mfc5_params.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_state orc_state;

orc_state* orc_create(const mfc5_params* P);
void orc_destroy(orc_state* S);
int orc_sys_size(const orc_state* S);
int orc_buff_size(const orc_state* S);

/* dir = 0,1,2.  s_cb: interior cell boundaries, length s+2 (index -1..s).  ghost_beg / ghost_end:
 * cell widths a neighbouring rank would deliver (length buff_size, ordered by increasing index),
 * used only when bc[dir][side] >= 0; NULL otherwise.  Restates m_boundary_common.fpp:404-466 and
 * m_mpi_common.fpp:1745-1796; then builds the WENO / CBC tables (m_weno.fpp:170-893,
 * m_cbc.fpp:327-408). */
void orc_set_grid(orc_state* S, int dir, const double* s_cb, const double* ghost_beg, const double* ghost_end);
void orc_finalize_setup(orc_state* S);
/* Dirichlet (-17) face values, layout [v][t2][t1] over the transverse interior (t1 < t2 the two
 * transverse directions); without it -17 degenerates to extrapolation (bc_io = F). */
void orc_set_dirichlet(orc_state* S, int dir, int side, const double* buf);

/* interior-packed (x fastest) N x (m+1)(n+1)(p+1) */
void orc_set_cons(orc_state* S, const double* q);
void orc_get_cons(const orc_state* S, double* q);
void orc_get_prim(const orc_state* S, double* q);
void orc_get_rhs(const orc_state* S, double* q);
/* the buffered boxes sf(-b:m+b,-b:n+b,-b:p+b) per variable (collapsed dims 0:0): which = 0 cons, 1 prim, 2 rhs */
long orc_box_size(const orc_state* S);
void orc_get_box(const orc_state* S, int which, double* q);
void orc_set_cons_box(orc_state* S, const double* q);

/* phases of one RK stage (m_rhs.fpp:574-902, m_time_steppers.fpp:458-507) */
void orc_cons2prim(orc_state* S);
void orc_bc_fill(orc_state* S, int dir, int side);          /* physical BCs only (bc < 0) */
int orc_halo_count(const orc_state* S, int dir);            /* doubles in one packed face slab */
void orc_halo_pack(const orc_state* S, int dir, int side, double* buf);   /* interior layers next to `side` */
void orc_halo_unpack(orc_state* S, int dir, int side, const double* buf); /* into the ghost layers of `side` */
void orc_sweeps(orc_state* S);                               /* WENO + Riemann + CBC + flux difference */
void orc_rk_update(orc_state* S, int stage, double dt);

/* single-rank convenience: one full RK stage / step with physical BCs on every side */
void orc_compute_rhs(orc_state* S);
void orc_tvd_rk_step(orc_state* S, double dt);
void orc_set_time(orc_state* S, double mytime);

/* max over interior cells of dt / min_d(ds_d / (|u_d| + c)) (m_sim_helpers.fpp:108-180) */
double orc_icfl_max(orc_state* S, double dt);

/* timing hook for bench.py's cpu baseline: OpenMP threads used by the sweeps (1 if built without) */
int orc_num_threads(void);
/* general WENO table construction (k = 3) against the transcribed WENO5 closed forms of direction dir: largest deviation */
double orc_selftest_weno_generic(orc_state* S, int dir);

#ifdef __cplusplus
}
#endif
#endif
