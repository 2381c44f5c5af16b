/* abi_driver.c -- drives libmfc5.so through include/mfc5.h exactly as INTEGRATION.md's Fortran shim does
 * (s_initialize_mfc5_b200 -> s_mfc5_upload -> s_tvd_rk ... -> mfc5_download_cons / mfc5_download_prim), from plain C:
 * no Python, no ctypes, the buffered q_cons_vf(i)%sf(-b:m+b,-b:n+b,-b:p+b) layout (ghosted = 1) on both sides.
 *
 *   abi_driver <in.bin> <out.bin>
 *   in.bin : mfc5_params | int nsteps | int use_n (1: mfc5_tvd_rk_n) | double dt | per direction: int len, len doubles x_cb(-1:s)
 *            | sys_size boxes of doubles (the caller's q_cons_vf, ghost cells holding NaN on purpose)
 *   out.bin: sys_size boxes q_cons_vf after the steps (ghost cells as the caller left them)
 *            | sys_size boxes q_prim_vf (ghost cells filled by the library) | double icfl | double stage_time_min
 * Built by tests/test_gpu_abi.py with `gcc -I include tests/c_driver/abi_driver.c -L mfc_b200 -lmfc5`.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mfc5.h"

#define DIE(msg) do { fprintf(stderr, "abi_driver: %s: %s\n", msg, mfc5_last_error()); return 2; } while (0)

static int rd(void* p, size_t n, FILE* f) { return fread(p, 1, n, f) == n ? 0 : -1; }

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: abi_driver in.bin out.bin\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    mfc5_params prm;
    int nsteps, use_n;
    double dt;
    if (rd(&prm, sizeof prm, f) || rd(&nsteps, sizeof nsteps, f) || rd(&use_n, sizeof use_n, f) || rd(&dt, sizeof dt, f)) return 1;

    mfc5_handle* h = NULL;
    if (mfc5_create(&prm, &h) != 0) DIE("mfc5_create");
    const int nd = 1 + (prm.n > 0) + (prm.p > 0);
    for (int d = 0; d < nd; d++) {   /* x_cb(-1:m), y_cb(-1:n), z_cb(-1:p); single rank: no neighbour widths */
        int len;
        if (rd(&len, sizeof len, f)) return 1;
        double* cb = (double*)malloc(sizeof(double) * (size_t)len);
        if (rd(cb, sizeof(double) * (size_t)len, f)) return 1;
        if (mfc5_set_grid(h, d, cb, NULL, NULL) != 0) DIE("mfc5_set_grid");
        free(cb);
    }
    if (mfc5_finalize_setup(h) != 0) DIE("mfc5_finalize_setup");

    const int N = mfc5_sys_size(h), b = mfc5_buff_size(h);
    const size_t ex = (size_t)prm.m + 1 + 2 * (size_t)b, ey = (size_t)prm.n + 1 + (prm.n > 0 ? 2 * (size_t)b : 0),
                 ez = (size_t)prm.p + 1 + (prm.p > 0 ? 2 * (size_t)b : 0), box = ex * ey * ez;
    /* sys_size separately allocated fields, like q_cons_vf(i)%sf: NOT adjacent in memory */
    double** q = (double**)malloc(sizeof(double*) * (size_t)N);
    double** pr = (double**)malloc(sizeof(double*) * (size_t)N);
    for (int v = 0; v < N; v++) {
        q[v] = (double*)malloc(sizeof(double) * box + 64 * (size_t)(v + 1));
        pr[v] = (double*)calloc(box, sizeof(double));
        if (rd(q[v], sizeof(double) * box, f)) return 1;
    }
    fclose(f);

    if (mfc5_upload_cons(h, (const double* const*)q, 1) != 0) DIE("mfc5_upload_cons");
    double time_avg = 0.0;
    if (use_n) {
        if (mfc5_tvd_rk_n(h, dt, nsteps) != 0) DIE("mfc5_tvd_rk_n");
        if (mfc5_sync(h) != 0) DIE("mfc5_sync");
    } else {
        for (int t = 0; t < nsteps; t++)
            if (mfc5_tvd_rk(h, dt, &time_avg) != 0) DIE("mfc5_tvd_rk");
    }
    double icfl = 0.0;
    if (mfc5_reduce(h, 0, dt, &icfl) != 0) DIE("mfc5_reduce");
    if (mfc5_download_cons(h, q, 1) != 0) DIE("mfc5_download_cons");
    if (mfc5_download_prim(h, pr, 1) != 0) DIE("mfc5_download_prim");

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    for (int v = 0; v < N; v++) fwrite(q[v], sizeof(double), box, f);
    for (int v = 0; v < N; v++) fwrite(pr[v], sizeof(double), box, f);
    fwrite(&icfl, sizeof icfl, 1, f);
    fwrite(&time_avg, sizeof time_avg, 1, f);
    fclose(f);
    printf("abi_driver: %d steps, sys_size %d, buff_size %d, %lld kernel launches, graph state %d\n", nsteps, N, b,
           mfc5_kernel_launches(h), mfc5_graph_state(h));
    if (mfc5_destroy(h) != 0) DIE("mfc5_destroy");
    return 0;
}
