/* mfc5.h -- C ABI of libmfc5.so, the B200-native (sm_100a CUDA, FP64) implementation of MFC's
 * 5-equation RHS hot path.  This is the drop-in boundary: the entry points are what a thin Fortran
 * ISO_C_BINDING layer binds in place of the reference's module procedures (INTEGRATION.md shows the
 * bind(C) interface block and the call sites).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error; mfc5_last_error() gives the message
 *     (the Fortran shim maps non-zero to s_mpi_abort, reference src/common/m_mpi_common.fpp:495).
 *   - a handle is single-caller (not thread safe), one handle per rank / per GPU, like the
 *     reference's one-rank-one-GPU model.
 *   - host field arguments are `sys_size` pointers, one per equation, exactly like
 *     q_cons_vf(1:sys_size)%sf (reference src/common/m_derived_types.fpp:21-24).  `ghosted` = 1:
 *     each array has the reference's buffered bounds (-b:m+b, -b:n+b, -b:p+b) (collapsed dims 0:0),
 *     x fastest; `ghosted` = 0: interior only (0:m,0:n,0:p).  Uploads read the interior only.  Downloads with
 *     `ghosted` = 1 write the whole array like the reference's GPU_UPDATE(host=...): q_prim_vf gets its ghost cells
 *     filled (faces, edges, corners, x -> y -> z order), q_cons_vf / rhs_vf get zeros in theirs (never meaningful in
 *     the reference either: boundary conditions act on the primitive variables, m_rhs.fpp:641-645).
 *   - all device memory is owned by the library; state persists in the handle between calls so a
 *     steady-state step moves no bytes between host and device.
 */
#ifndef MFC5_H
#define MFC5_H
#include "mfc5_params.h"

/* the library is built with -fvisibility=hidden; only the C ABI below is exported */
#if defined(__GNUC__)
#define MFC5_API __attribute__((visibility("default")))
#else
#define MFC5_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfc5_handle mfc5_handle;

/* replaces s_initialize_{weno,riemann_solvers,cbc,rhs,time_steppers}_module
 * (reference src/simulation/m_start_up.fpp:840-946): validates parameters (m_checker.fpp:52-69),
 * allocates device state. */
MFC5_API int mfc5_create(const mfc5_params* p, mfc5_handle** out);

/* replaces s_finalize_*_module (m_start_up.fpp:1106-1141) */
MFC5_API int mfc5_destroy(mfc5_handle* h);

/* Grid of one direction (0,1,2).  s_cb: interior cell boundaries x_cb(-1:m) (m_start_up.fpp:179-191).
 * Ghost metrics are rebuilt as s_populate_grid_variables_buffers does (m_boundary_common.fpp:384-468);
 * on processor boundaries (bc >= 0) the caller passes the `buff_size` neighbour cell widths that
 * s_mpi_sendrecv_grid_variable_buffer would deliver (m_mpi_common.fpp:1745-1796), else NULL.
 * Call once per direction, then mfc5_finalize_setup (builds WENO / CBC tables:
 * s_compute_weno_coefficients m_weno.fpp:170-893, s_compute_cbc_coefficients m_cbc.fpp:327-408). */
MFC5_API int mfc5_set_grid(mfc5_handle* h, int dir, const double* s_cb, const double* ghost_ds_beg, const double* ghost_ds_end);
MFC5_API int mfc5_finalize_setup(mfc5_handle* h);

/* Dirichlet (-17) face values = the reference's bc_buffers(dir,loc)%sf read by
 * s_read_serial_boundary_condition_files (src/common/m_boundary_io.fpp:160-205).
 * Layout [sys_size][t2][t1] over the transverse interior. */
MFC5_API int mfc5_set_dirichlet(mfc5_handle* h, int dir, int side, const double* values);

/* replaces the GPU_UPDATE(device=...) of q_cons_ts(1) in s_initialize_gpu_vars (m_start_up.fpp:1034-1042) */
MFC5_API int mfc5_upload_cons(mfc5_handle* h, const double* const* q_cons_vf, int ghosted);
/* replaces GPU_UPDATE(host=...) in s_save_data (m_start_up.fpp:739) */
MFC5_API int mfc5_download_cons(mfc5_handle* h, double* const* q_cons_vf, int ghosted);
/* q_prim_vf of the last RHS evaluation, ghost cells filled (m_rhs.fpp:884-898); ghosted layout fills
 * the buffer cells too */
MFC5_API int mfc5_download_prim(mfc5_handle* h, double* const* q_prim_vf, int ghosted);
/* rhs_vf(1:sys_size)%sf(0:m,0:n,0:p) of the last mfc5_compute_rhs */
MFC5_API int mfc5_download_rhs(mfc5_handle* h, double* const* rhs_vf, int ghosted);

/* replaces s_compute_rhs(q_cons_vf, ..., rhs_vf, ..., t_step, stage) (m_rhs.fpp:574): evaluates
 * the RHS of the current conservative state into the device rhs buffer (no state update). */
MFC5_API int mfc5_compute_rhs(mfc5_handle* h);

/* replaces s_tvd_rk(t_step, time_avg, nstage) (m_time_steppers.fpp:438-611): one full TVD-RK step
 * of size dt (time_stepper stages, each = RHS + halo + fused RK update).  stage_time_min, if not
 * NULL, is updated with the reference's grind timer rule (minimum stage wall time in seconds,
 * m_time_steppers.fpp:559-577): pass the running value (<= 0 to start). */
MFC5_API int mfc5_tvd_rk(mfc5_handle* h, double dt, double* stage_time_min);

/* Enqueue `nsteps` steps of size dt without host synchronisation between them (CUDA-graph replay of
 * the captured step); returns after enqueue.  mfc5_sync waits. */
MFC5_API int mfc5_tvd_rk_n(mfc5_handle* h, double dt, int nsteps);
MFC5_API int mfc5_sync(mfc5_handle* h);
/* how mfc5_tvd_rk_n runs its steps: 1 = CUDA-graph replay, 0 = eager launches by request, -1 = a graph capture failed and
 * the handle fell back to eager launches (a performance cliff on small grids, never a correctness issue) */
MFC5_API int mfc5_graph_state(const mfc5_handle* h);

/* Simulation time of the state (the reference's `mytime`, GPU_UPDATE(device=[mytime]) m_time_steppers.fpp:532); only the
 * body-force acceleration g + k sin(w t - p) depends on it.  mfc5_tvd_rk / mfc5_tvd_rk_n add dt after every step, like
 * s_perform_time_step (m_start_up.fpp:634). */
MFC5_API int mfc5_set_time(mfc5_handle* h, double mytime);

/* reductions over the interior, using the primitive variables of the current state:
 * what = 0: ICFL max  = max dt/min_d(ds_d/(|u_d|+c))  (m_sim_helpers.fpp:108-130, m_data_output.fpp:159-224)
 * what = 1: dt min    = min_d(ds_d/(|u_d|+c)) (caller multiplies by cfl_target; m_sim_helpers.fpp:182-249)
 * what = 2: NaN count of the conservative state (the reference's save-time NaN scan, m_start_up.fpp:741-751)
 * In a multi-rank run the result is reduced over ranks (ncclAllReduce), replacing
 * s_mpi_allreduce_min / s_mpi_reduce_stability_criteria_extrema (m_mpi_common.fpp:330-342,439). */
MFC5_API int mfc5_reduce(mfc5_handle* h, int what, double dt, double* out);

/* Multi-GPU: replaces MPI_CART / s_mpi_sendrecv_variables_buffers (m_mpi_common.fpp:549-1070) with
 * packed halo faces exchanged by ncclSend/ncclRecv on a communication stream.  nccl_unique_id is the
 * 128-byte ncclUniqueId created by rank 0 (mfc5_nccl_unique_id) and broadcast by the host
 * (torch.distributed / MPI_Bcast).  Neighbour ranks are the non-negative bc[][] entries. */
MFC5_API int mfc5_nccl_unique_id(char id_out[128]);
MFC5_API int mfc5_comm_init(mfc5_handle* h, const char nccl_unique_id[128], int rank, int nranks);

/* introspection used by tests and bench.py */
MFC5_API int mfc5_sys_size(const mfc5_handle* h);
MFC5_API int mfc5_buff_size(const mfc5_handle* h);
MFC5_API long long mfc5_kernel_launches(const mfc5_handle* h);   /* library kernels launched so far */
/* device time (ms, CUDA events on the library's compute stream) and launch count of one kernel class,
 * summed over the stages of the last mfc5_tvd_rk call; returns < 0 for a bad argument.  The classes
 * are the analogue of the reference's NVTX ranges (m_rhs.fpp:594,640,644,912,1030,1041). */
enum {
    MFC5_K_CONS2PRIM = 0, /* RHS-CONVERT */
    MFC5_K_BC_FILL = 1,   /* RHS-COMMUNICATION, physical boundaries */
    MFC5_K_HALO = 2,      /* RHS-COMMUNICATION, pack + ncclSend/Recv + unpack */
    MFC5_K_CBC = 3,       /* characteristic-BC planes */
    MFC5_K_SWEEP_X = 4,   /* RHS-RECONSTRUCTION + RIEMANN-SOLVER + ADVECTION-SRC, fused, per direction */
    MFC5_K_SWEEP_Y = 5,
    MFC5_K_SWEEP_Z = 6,   /* the last direction of a stage also carries the RK update */
    MFC5_NUM_KCLASS = 7
};
MFC5_API double mfc5_last_kernel_ms(const mfc5_handle* h, int kclass, long long* launches);
MFC5_API void* mfc5_compute_stream(const mfc5_handle* h);        /* cudaStream_t */
MFC5_API const char* mfc5_last_error(void);
MFC5_API const char* mfc5_version(void);

#ifdef __cplusplus
}
#endif
#endif
