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
    runlat<0>("DFMA", 1);
    runlat<2>("DMUL", 1);
    runlat<1>("MUFU.RCP64H+DADD", 1);
    return 0;
}
