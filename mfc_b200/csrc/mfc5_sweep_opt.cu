// "case-optimised" fused sweep kernels: compile-time WENO order / variant / Riemann solver for the
// configurations of BASELINE.json (3-D, two fluids).  The analogue of the reference's
// --case-optimization (docs/documentation/expectedPerformance.md:11-15).
#include "mfc5_host.cuh"
namespace mfc5 {
bool launch_sweep_opt(const DevP& P, int dir, cudaStream_t st) {
#ifdef MFC5_STRICT
    return false;
#else
    if (!opt_config(P)) return false;
    if (P.rs == 2) {
        if (P.wo == 5 && P.wv == 1) return launch_sweep_opt_t<2, 3, 5, 1, 2>(P, dir, st);  // 3D_performance_test (WENO5-M, HLLC)
        if (P.wo == 5 && P.wv == 2) return launch_sweep_opt_t<2, 3, 5, 2, 2>(P, dir, st);  // WENO5-Z override of 3D_shockdroplet
        if (P.wo == 3 && P.wv == 1) return launch_sweep_opt_t<2, 3, 3, 1, 2>(P, dir, st);  // benchmarks/5eq_rk3_weno3_hllc, 3D_shockdroplet
    } else if (P.rs == 1) {
        if (P.wo == 5 && P.wv == 1) return launch_sweep_opt_t<2, 3, 5, 1, 1>(P, dir, st);  // 3D_performance_test HLL variant
    }
    return false;
#endif
}
}  // namespace mfc5
