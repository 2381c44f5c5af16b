// mfc5_host.cuh -- internal interface between the C-ABI host code and the sweep translation units.
#pragma once
#include "mfc5_kernels.cuh"

namespace mfc5 {

// Launch the fused sweep of direction `dir` for the shape / scheme in P on `st`.
// Returns false if this translation unit has no kernel for P (caller tries the next one).
// Specialised ("case-optimised") kernels: compile-time WENO order / variant / Riemann solver.
bool launch_sweep_opt(const DevP& P, int dir, cudaStream_t st);
// Generic kernels: run-time scheme selection, every supported <NF, ND>.
bool launch_sweep_generic_nf1(const DevP& P, int dir, cudaStream_t st);
bool launch_sweep_generic_nf2(const DevP& P, int dir, cudaStream_t st);
bool launch_sweep_generic_nf3(const DevP& P, int dir, cudaStream_t st);

inline void sweep_grid(const DevP& P, int dir, dim3& grid, dim3& block) {
    block = dim3(128, 1, 1);
    if (dir == 0) {
        const long long segs = (P.s[0] + 1 + MFC5_XW - 1) / MFC5_XW;
        const long long warps = segs * (long long)(P.s[1] + 1) * (P.s[2] + 1);
        grid = dim3((unsigned)((warps * 32 + 127) / 128), 1, 1);
    } else {
        const long long plane = (long long)(P.s[0] + 1) * ((dir == 1 ? P.s[2] : P.s[1]) + 1);
        const int chunks = (P.s[dir] + 1 + P.chunk - 1) / P.chunk;
        grid = dim3((unsigned)((plane + 127) / 128), (unsigned)chunks, 1);
    }
}

template <int NF, int ND, int WO, int WV, int RS, int FAST>
inline bool launch_sweep_t(const DevP& P, int dir, cudaStream_t st) {
    dim3 grid, block;
    sweep_grid(P, dir, grid, block);
    if (dir == 0) {
        k_sweep_x<NF, ND, WO, WV, RS, FAST><<<grid, block, 0, st>>>(P);
        return true;
    }
    if constexpr (ND >= 2) {
        if (dir == 1) {
            k_sweep_yz<NF, ND, 1, WO, WV, RS, FAST><<<grid, block, 0, st>>>(P);
            return true;
        }
    }
    if constexpr (ND >= 3) {
        if (dir == 2) {
            k_sweep_yz<NF, ND, 2, WO, WV, RS, FAST><<<grid, block, 0, st>>>(P);
            return true;
        }
    }
    return false;
}

// P.fast selects the division-light arithmetic (product default) or the reference operation order
// (bitwise comparable with the CPU oracle when built with -fmad=false).
template <int NF>
inline bool launch_sweep_generic(const DevP& P, int dir, cudaStream_t st) {
    if (P.nf != NF) return false;
#ifndef MFC5_STRICT
    if (P.fast) {
        switch (P.nd) {
        case 1: return launch_sweep_t<NF, 1, -1, -1, -1, 1>(P, dir, st);
        case 2: return launch_sweep_t<NF, 2, -1, -1, -1, 1>(P, dir, st);
        case 3: return launch_sweep_t<NF, 3, -1, -1, -1, 1>(P, dir, st);
        }
    }
#endif
    switch (P.nd) {
    case 1: return launch_sweep_t<NF, 1, -1, -1, -1, 0>(P, dir, st);
    case 2: return launch_sweep_t<NF, 2, -1, -1, -1, 0>(P, dir, st);
    case 3: return launch_sweep_t<NF, 3, -1, -1, -1, 0>(P, dir, st);
    }
    return false;
}

}  // namespace mfc5
