// fp64_peak.cu -- measures the FP64 vector pipe of the GPU it runs on: DFMA / DMUL / DADD throughput and
// the cost of the reciprocal sequence used by the division-light kernels (MUFU.RCP64H + 4 DFMA) and of
// an IEEE division.  The result is the denominator of the "FP64 pipe utilisation" figures in DESIGN.md /
// profiles/ (MEASURED_PEAKS.json has no FP64 entry).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + 1e-3 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = fma(a[i], b, c);
                else if (OP == 1) a[i] = a[i] * b;
                else if (OP == 2) a[i] = a[i] + c;
                else if (OP == 3) {  // rcp_nr
                    double r;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i]));
                    double e = fma(-a[i], r, 1.0);
                    r = fma(r, e, r);
                    e = fma(-a[i], r, 1.0);
                    a[i] = fma(r, e, r) + 1.5;
                } else if (OP == 4) a[i] = 1.0 / a[i] + 1.5;
                else if (OP == 5) a[i] = sqrt(a[i]) + 1.5;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP>
double run(const char* name, int nsm, double ops_per_elem, int clock_khz) {
    double* d;
    cudaMalloc(&d, 8);
    const int blocks = nsm * 8, iters = 2048;
    k<OP><<<blocks, 256>>>(d, 16, 1.0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(d, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double n = (double)blocks * 256 * iters * 32.0;  // element-ops
    const double rate = n / (best * 1e-3);
    printf("%-10s %8.3f ms  %8.2f G elem-op/s  = %6.2f per clk per SM at %d MHz  (%.2f TFLOP/s if FMA)\n", name, best, rate / 1e9,
           rate / nsm / (clock_khz * 1e3), clock_khz / 1000, rate * ops_per_elem / 1e12);
    cudaFree(d);
    return rate;
}

// dependent-issue latency: one warp, one chain
template <int OP>
__global__ void lat(double* out, long long* cyc, double seed) {
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) a = fma(a, b, c);
            else if (OP == 1) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 1.5; }
            else if (OP == 2) a = a * b;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = a; cyc[0] = t1 - t0; }
}
template <int OP>
void runlat(const char* name, int chain_ops) {
    double* d; long long* c; long long h;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    lat<OP><<<1, 32>>>(d, c, 1.0);
    lat<OP><<<1, 32>>>(d, c, 1.0);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("latency %-14s %.2f cycles per dependent step (%d FP64 ops per step)\n", name, (double)h / (256.0 * 16.0), chain_ops);
    cudaFree(d); cudaFree(c);
}

// DFMA with three distinct register operands per instruction (operand-bandwidth / bank-conflict check)
__global__ void __launch_bounds__(256) k3(double* out, int iters, double seed) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + 1e-3 * (threadIdx.x + i); b[i] = 1.0 + 1e-7 * (i + 1); c[i] = 1e-9 * (i + 1); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b[(i + u) & 7], c[(i + 2 * u + 1) & 7]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i] + c[i];
    if (s == 12345.678) out[0] = s;
}
// 4 independent chains only (ILP 4) at 8, 12, 16 warps per SM: how many warps does the pipe need?
template <int ILP>
__global__ void kilp(double* out, int iters, double seed) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = seed + 1e-3 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 32 / ILP; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}
// DFMA mixed with integer / shared-memory instructions: K non-FP64 instructions per 8 DFMA (the sweeps run
// at about 5 of 8), 12 warps per SM, 4 independent DFMA chains per thread
template <int K>
__global__ void kmix(double* out, int iters, double seed, int* sink) {
    __shared__ int sm[384];
    double a[4];
    int x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = seed + 1e-3 * (threadIdx.x + i); x[i] = threadIdx.x + i; }
    sm[threadIdx.x] = threadIdx.x;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = fma(a[i], b, c);
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = fma(a[i], b, c);
#pragma unroll
            for (int k = 0; k < K; k++) x[k & 3] = x[k & 3] * 3 + (x[(k + 1) & 3] ^ it);
        }
    }
    double s = 0; int y = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { s += a[i]; y += x[i]; }
    if (s == 12345.678) out[0] = s;
    if (y == 123456789) sink[0] = y;
}
template <int K>
void runmix(int nsm, int khz) {
    double* d; int* sk; cudaMalloc(&d, 8); cudaMalloc(&sk, 4);
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kmix<K><<<nsm, 384>>>(d, iters, 1.0, sk);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rate = (double)nsm * 384 * iters * 64.0 / (best * 1e-3);
    printf("mix: %d integer instr per 8 DFMA, 12 warps/SM: %6.2f DFMA per clk per SM (of 64)\n", K, rate / nsm / (khz * 1e3));
    cudaFree(d); cudaFree(sk);
}

template <int ILP>
void runilp(int nsm, int warps_per_sm, int khz) {
    double* d; cudaMalloc(&d, 8);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kilp<ILP><<<nsm, warps_per_sm * 32>>>(d, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rate = (double)nsm * warps_per_sm * 32 * iters * 32.0 / (best * 1e-3);
    printf("ILP %d, %2d warps/SM: %6.2f DFMA per clk per SM (of 64)\n", ILP, warps_per_sm, rate / nsm / (khz * 1e3));
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, khz / 1000);
    const int nsm = p.multiProcessorCount;
    run<0>("DFMA", nsm, 2, khz);
    run<1>("DMUL", nsm, 1, khz);
    run<2>("DADD", nsm, 1, khz);
    run<3>("rcp_nr+add", nsm, 1, khz);
    run<4>("1/x+add", nsm, 1, khz);
    run<5>("sqrt+add", nsm, 1, khz);
    {
        double* d; cudaMalloc(&d, 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); k3<<<nsm * 8, 256>>>(d, 2048, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double rate = (double)nsm * 8 * 256 * 2048 * 32.0 / (best * 1e-3);
        printf("DFMA 3 distinct operands: %6.2f per clk per SM\n", rate / nsm / (khz * 1e3));
    }
    for (int w : {4, 8, 12, 16, 24}) { runilp<1>(nsm, w, khz); runilp<2>(nsm, w, khz); runilp<4>(nsm, w, khz); }
    runmix<0>(nsm, khz); runmix<2>(nsm, khz); runmix<4>(nsm, khz); runmix<5>(nsm, khz); runmix<6>(nsm, khz); runmix<8>(nsm, khz);
    runlat<0>("DFMA", 1);
    runlat<2>("DMUL", 1);
    runlat<1>("MUFU.RCP64H+DADD", 1);
    return 0;
}
