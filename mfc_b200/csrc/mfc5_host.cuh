// mfc5_host.cuh -- internal interface between the C-ABI host code and the sweep translation units.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "mfc5_kernels.cuh"

namespace mfc5 {

// Does the case-optimised path (mfc5_sweep_opt.cu) take this configuration?  The host asks the same question to decide
// whether ghost cells are folded into the sweep loaders (DevP::gmode) and the cons -> prim pass into the last sweep.
inline bool opt_config(const DevP& P) {
#ifdef MFC5_STRICT
    return false;
#else
    if (getenv("MFC5_NO_OPT")) return false;
    if (P.nf != 2 || P.nd != 3 || !P.fast || !P.fast_riem) return false;
    if ((unsigned long long)P.fs * (unsigned long long)(2 * P.nf + P.nd + 1) >= (1ull << 32)) return false;  // 32-bit field indices
    for (int d = 0; d < 3; d++)
        for (int sd = 0; sd < 2; sd++)
            if (P.bc[d][sd] == MFC5_BC_RIEMANN_EXTRAP) return false;   // BC -4 lives in the generic kernels
    if (P.rs == 2) return (P.wo == 5 && (P.wv == 1 || P.wv == 2)) || (P.wo == 3 && P.wv == 1);
    if (P.rs == 1) return P.wo == 5 && P.wv == 1;
    return false;
#endif
}
// ... and are its y / z sweeps the second-generation marching kernel (which is the one that implements gmode and prim_next)?
inline bool opt_march2(const DevP& P) {
    return opt_config(P) && !getenv("MFC5_MARCH_V1");
}

// Launch the fused sweep of direction `dir` for the shape / scheme in P on `st`.
// Returns false if this translation unit has no kernel for P (caller tries the next one).
// Specialised ("case-optimised") kernels: compile-time WENO order / variant / Riemann solver.
bool launch_sweep_opt(const DevP& P, int dir, cudaStream_t st);
// Generic kernels: run-time scheme selection, every supported <NF, ND>.
bool launch_sweep_generic_nf1(const DevP& P, int dir, cudaStream_t st);
bool launch_sweep_generic_nf2(const DevP& P, int dir, cudaStream_t st);
bool launch_sweep_generic_nf3(const DevP& P, int dir, cudaStream_t st);

// x-sweep, warp-shuffle kernel: used when there are too few pencils for the marching kernel (1-D, thin 2-D)
inline void sweep_grid_x(const DevP& P, dim3& grid, dim3& block) {
    const long long segs = P.xseg_n > 0 ? P.xseg_n : (P.s[0] + 1 + MFC5_XW - 1) / MFC5_XW;
    const int ny = P.s[1] + 1, nz = P.s[2] + 1;
    const int nw = (ny >= 4) ? 4 : 1;  // warps per block (k_sweep_x): 4 rows of one segment at a time, 1 for thin domains
    block = dim3(32 * nw, 1, 1);
    const int rows_per_block = MFC5_X_RPB * nw;
    grid = dim3((unsigned)segs, (unsigned)((ny + rows_per_block - 1) / rows_per_block), (unsigned)nz);
}

inline long long march_pencils(const DevP& P, int dir) {
    return dir == 0 ? (long long)(P.s[1] + 1) * (P.s[2] + 1) : (long long)(P.s[0] + 1) * ((dir == 1 ? P.s[2] : P.s[1]) + 1);
}

// chunk length along the marching direction: the number of blocks should be close to a whole number of
// waves of the resident-block capacity (148 SMs x blocks/SM), chunks as long as possible (each chunk
// re-does 2 cells of reconstruction and a (2*PN)-plane prologue).
inline int march_chunk_n(int n, long long lanes_blocks, int blocks_per_sm, int max_chunk, int extra) {
    const double slots = 148.0 * blocks_per_sm;
    int best = n;
    double best_cost = 1e300;
    for (int nch = 1; nch <= n; nch++) {
        const int chunk = (n + nch - 1) / nch;
        if (chunk > max_chunk) continue;
        if (chunk < 4 && nch > 1) break;
        const int real = (n + chunk - 1) / chunk;
        const double waves = std::ceil(lanes_blocks * real / slots);
        const double cost = waves * (chunk + extra + 1.0);  // per-block time ~ iterations + prologue
        if (cost < best_cost) { best_cost = cost; best = chunk; }
    }
    return best;
}
inline int march_chunk(const DevP& P, int dir, int blocks_per_sm, int max_chunk) {
    return march_chunk_n(P.s[dir] + 1, (march_pencils(P, dir) + MFC5_MARCH_THREADS - 1) / MFC5_MARCH_THREADS, blocks_per_sm, max_chunk, 2);
}

template <int NF, int ND, int DIR, int WO, int WV, int RS, int FAST>
inline bool launch_march(const DevP& P0, cudaStream_t st) {
    // resident blocks per SM the kernel is built for; one when two rings do not fit (order 7 with 8 variables)
    const int bps = ((2 * NF + ND + 1 <= 8) && 2 * (march_smem_bytes<NF, ND, WO>(4) + 1024) <= 227 * 1024) ? MFC5_MARCH_MINB : 1;
    auto kern = k_sweep_march<NF, ND, DIR, WO, WV, RS, FAST>;
    // the attribute is per device (a process may hold handles on several): once per (kernel, device ordinal)
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    DevP P = P0;
    // longest chunk whose table rows still let `bps` blocks share the 227 KB of an SM
    const long long budget = (227LL * 1024 - bps * 1024) / bps - (long long)march_smem_bytes<NF, ND, WO>(-2);
    const int max_chunk = (int)std::max<long long>(4, budget / (MFC5_TAB * (long long)sizeof(double)) - 2);
    if (march_smem_bytes<NF, ND, WO>(4) > 226 * 1024) return false;   // does not fit at all
    if (P.chunk <= 0) P.chunk = march_chunk(P, DIR, bps, max_chunk);
    P.chunk = std::min(P.chunk, max_chunk);
    const long long lanes_blocks = (march_pencils(P, DIR) + MFC5_MARCH_THREADS - 1) / MFC5_MARCH_THREADS;
    const int chunks = (P.s[DIR] + 1 + P.chunk - 1) / P.chunk;
    kern<<<dim3((unsigned)lanes_blocks, (unsigned)chunks, 1), MFC5_MARCH_THREADS, march_smem_bytes<NF, ND, WO>(P.chunk), st>>>(P);
    return true;
}

template <int NF, int ND, int WO, int WV, int RS, int FAST>
inline bool launch_sweep_t(const DevP& P, int dir, cudaStream_t st) {
    if (dir == 0) {
        if (march_pencils(P, 0) < 4096 || !P.x_march) {
            dim3 grid, block;
            sweep_grid_x(P, grid, block);
            k_sweep_x<NF, ND, WO, WV, RS, FAST><<<grid, block, 0, st>>>(P);
            return true;
        }
        if constexpr (ND >= 2) return launch_march<NF, ND, 0, WO, WV, RS, FAST>(P, st);
    }
    if constexpr (ND >= 2) {
        if (dir == 1) return launch_march<NF, ND, 1, WO, WV, RS, FAST>(P, st);
    }
    if constexpr (ND >= 3) {
        if (dir == 2) return launch_march<NF, ND, 2, WO, WV, RS, FAST>(P, st);
    }
    return false;
}

// second-generation marching kernel (case-optimised configurations; k_march2)
template <int NF, int ND, int DIR, int WO, int WV, int RS>
inline bool launch_march2(const DevP& P0, cudaStream_t st) {
    constexpr int bps = MFC5_M2_MINB;
    auto kern = k_march2<NF, ND, DIR, WO, WV, RS>;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    DevP P = P0;
    // longest chunk whose table rows still let `bps` blocks share the 228 KB of an SM (1 KB per block is reserved by the system)
    const long long budget = (228LL * 1024 - bps * 1024) / bps - (long long)march2_smem_bytes<NF, ND, WO>(-2);
    int max_chunk = P.s[DIR] + 1;
    if (MFC5_M2_TABS) {
        if (budget < (long long)(6 * MFC5_TAB * sizeof(double))) return false;
        max_chunk = (int)(budget / (MFC5_TAB * (long long)sizeof(double)) - 2);
    }
    const long long lanes_blocks = (march_pencils(P, DIR) + MFC5_M2_THREADS - 1) / MFC5_M2_THREADS;
    if (P.chunk <= 0) P.chunk = march_chunk_n(P.s[DIR] + 1, lanes_blocks, bps, max_chunk, 2);
    P.chunk = std::min(P.chunk, max_chunk);
    const int chunks = (P.s[DIR] + 1 + P.chunk - 1) / P.chunk;
    kern<<<dim3((unsigned)lanes_blocks, (unsigned)chunks, 1), MFC5_M2_THREADS, march2_smem_bytes<NF, ND, WO>(P.chunk), st>>>(P);
    return true;
}

// case-optimised sweep: warp-shuffle kernel along x, second-generation marching kernel along y / z
template <int NF, int ND, int WO, int WV, int RS>
inline bool launch_sweep_opt_t(const DevP& P, int dir, cudaStream_t st) {
    if (dir == 0 || !opt_march2(P)) return launch_sweep_t<NF, ND, WO, WV, RS, 1>(P, dir, st);
    if constexpr (ND >= 2) {
        if (dir == 1) return launch_march2<NF, ND, 1, WO, WV, RS>(P, st);
    }
    if constexpr (ND >= 3) {
        if (dir == 2) return launch_march2<NF, ND, 2, WO, WV, RS>(P, st);
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
    if (P.wo == 7) {   // seven-point window: its own instantiation (reference-order arithmetic, run-time variant / solver)
        switch (P.nd) {
        case 1: return launch_sweep_t<NF, 1, 7, -1, -1, 0>(P, dir, st);
        case 2: return launch_sweep_t<NF, 2, 7, -1, -1, 0>(P, dir, st);
        case 3: return launch_sweep_t<NF, 3, 7, -1, -1, 0>(P, dir, st);
        }
    }
    switch (P.nd) {
    case 1: return launch_sweep_t<NF, 1, -1, -1, -1, 0>(P, dir, st);
    case 2: return launch_sweep_t<NF, 2, -1, -1, -1, 0>(P, dir, st);
    case 3: return launch_sweep_t<NF, 3, -1, -1, -1, 0>(P, dir, st);
    }
    return false;
}

}  // namespace mfc5
