// generic (run-time scheme) fused sweep kernels, num_fluids = 2
#include "mfc5_host.cuh"
namespace mfc5 {
bool launch_sweep_generic_nf2(const DevP& P, int dir, cudaStream_t st) { return launch_sweep_generic<2>(P, dir, st); }
}  // namespace mfc5
