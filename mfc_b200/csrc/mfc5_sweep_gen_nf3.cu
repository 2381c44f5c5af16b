// generic (run-time scheme) fused sweep kernels, num_fluids = 3
#include "mfc5_host.cuh"
namespace mfc5 {
bool launch_sweep_generic_nf3(const DevP& P, int dir, cudaStream_t st) { return launch_sweep_generic<3>(P, dir, st); }
}  // namespace mfc5
