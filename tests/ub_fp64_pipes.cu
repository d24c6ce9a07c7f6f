// Microbenchmark: do DMMA (mma.sync m8n8k4 f64) and DFMA share an execution pipe on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA; 3: only even warps (DMMA), odd idle; 4: only odd (DFMA)
__global__ void __launch_bounds__(512) k(int mode, int iters, double* out) {
    const int warp = threadIdx.x >> 5;
    const bool even = (warp & 1) == 0;
    const bool do_dmma = mode == 0 || ((mode == 2 || mode == 3) && even);
    const bool do_dfma = mode == 1 || ((mode == 2 || mode == 4) && !even);
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    if (do_dmma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
        }
    } else if (do_dfma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* d; cudaMalloc(&d, 8);
    const int iters = 20000, ctas = 148 * 2, threads = 512;
    const char* names[] = {"all DMMA", "all DFMA", "even DMMA + odd DFMA", "even DMMA only", "odd DFMA only"};
    for (int mode = 0; mode < 5; ++mode) {
        k<<<ctas, threads>>>(mode, 100, d); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<ctas, threads>>>(mode, iters, d); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)ctas * threads / 32;
        double w_dmma = mode == 0 ? warps : (mode == 2 || mode == 3) ? warps / 2 : 0;
        double w_dfma = mode == 1 ? warps : (mode == 2 || mode == 4) ? warps / 2 : 0;
        const double f_dmma = w_dmma * iters * 8.0 * 512.0, f_dfma = w_dfma * iters * 64.0 * 64.0;
        printf("%-24s %8.3f ms  DMMA %7.2f TF/s  DFMA %7.2f TF/s  total %7.2f TF/s\n", names[mode], ms, f_dmma / ms / 1e9, f_dfma / ms / 1e9,
               (f_dmma + f_dfma) / ms / 1e9);
    }
    return 0;
}
