// Microbenchmark: the inner loop of dgemm_dmma_kernel (fragment loads from padded shared memory + DMMA) WITHOUT any
// global traffic: what does the LDS + DMMA + barrier structure alone reach at 3 CTAs x 4 warps per SM?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int BK = 16, PADM = 68, PADK = 20, STAGES = 3;
constexpr int A_TILE = 64 * PADK, STAGE = 2 * A_TILE;      // A as [k][m] needs 16 x 68 = 1088 <= 1280
template <int MI, int NJ, bool BAR>
__global__ void __launch_bounds__(128, 3) k(int ktiles, double* out) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / 2, wn = warp % 2, lr = lane >> 2, lc = lane & 3;
    for (int i = tid; i < STAGES * STAGE; i += 128) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < ktiles; ++kt) {
        if (BAR) __syncthreads();
        const double* As = sm + (kt % STAGES) * STAGE;
        const double* Bs = As + A_TILE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[MI], bf[NJ];
            const int kq = kk * 4 + lc;
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = As[kq * PADM + (wm * MI * 8 + i * 8 + lr) % 64];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = Bs[((wn * NJ * 8 + j * 8 + lr) % 64) * PADK + kq];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}
template <int MI, int NJ, bool BAR>
void run(const char* name, double* d) {
    const int ktiles = 20000, ctas = 148 * 3;
    const size_t smem = STAGES * STAGE * sizeof(double);
    cudaFuncSetAttribute(k<MI, NJ, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MI, NJ, BAR><<<ctas, 128, smem>>>(100, d); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MI, NJ, BAR><<<ctas, 128, smem>>>(ktiles, d); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)ctas * 4 * ktiles * (BK / 4) * MI * NJ * 512.0;
    printf("%-44s %8.3f ms  %6.2f TFLOP/s  (%s)\n", name, ms, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double* d; cudaMalloc(&d, 8);
    run<4, 4, true>("warp 32x32, barrier per k-tile (the kernel)", d);
    run<4, 4, false>("warp 32x32, no barrier", d);
    run<4, 8, true>("warp 32x64, barrier per k-tile", d);
    run<4, 8, false>("warp 32x64, no barrier", d);
    run<2, 4, true>("warp 16x32, barrier per k-tile", d);
    return 0;
}
