// Host-matrix row-chunk pipelines shared by the single-device (api.cu) and multi-device (multi.cu) entry points.
#pragma once
#include "kb_common.cuh"

namespace kb {

// column-major host matrix (rows x cols, ld hld) <-> device (ld dld)
static inline int h2d_matrix(kb_ctx* ctx, const double* h, long hld, double* d, long dld, long rows, long cols, cudaStream_t st) {
    KB_CUDA(ctx, cudaMemcpy2DAsync(d, dld * sizeof(double), h, hld * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyHostToDevice, st));
    return KB_OK;
}
static inline int d2h_matrix(kb_ctx* ctx, const double* d, long dld, double* h, long hld, long rows, long cols, cudaStream_t st) {
    KB_CUDA(ctx, cudaMemcpy2DAsync(h, hld * sizeof(double), d, dld * sizeof(double), rows * sizeof(double), cols,
                                   cudaMemcpyDeviceToHost, st));
    return KB_OK;
}

// true if `ptr` is ordinary pageable host memory (neither cudaHostAlloc'ed nor cudaHostRegister'ed nor managed)
bool host_is_pageable(const void* ptr);
// pinned bounce buffer `idx` of ctx with at least `bytes` bytes (grown on demand, kept between calls)
int ensure_pinned(kb_ctx* ctx, int idx, size_t bytes, double** out);
void release_pinned(kb_ctx* ctx);
// dst(ld dld) <- src(ld sld), rows x cols doubles, split by columns over a few host threads
void parallel_copy2d(double* dst, long dld, const double* src, long sld, long rows, long cols);

// Declared first in a function that enqueues work on ctx's three pipeline streams: whatever way the function is left
// (KB_TRY / KB_CUDA early returns included), every stream is drained BEFORE the Scope declared above it hands the device
// slabs back to the pool and before the caller may release the host buffers the copies read or write.
struct StreamDrain {
    kb_ctx* ctx;
    explicit StreamDrain(kb_ctx* c) : ctx(c) {}
    ~StreamDrain() {
        cudaStreamSynchronize(ctx->stream_h2d);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream_d2h);
    }
};

// Marginal-statistic accumulators fed by the sampling pipeline chunk by chunk (fit_marginal: X̃ is consumed while it
// is still on the device).
struct StatsHook {
    const double* dys;   // y_std, n doubles on the device
    double* dacc_x;      // STATS_ACC * p
    double* dacc_ko;     // STATS_ACC * m p
};

// Row-chunk sampling pipeline for host-resident X / Z / X̃ (api.cu).  hld: leading dimension of the host matrices
// (0 or < n: n), so a row shard of larger matrices can be passed as (X + lo, n_local, hld = n_total).
int sample_host_pipeline(kb_ctx* ctx, const double* X, long n, int p, const double* dmu, const double* dSiS,
                         const double* dLb, long ldo, int m, const double* Z, uint64_t seed, int64_t row_offset,
                         double* Xko, const StatsHook* hook = nullptr, long hld = 0);

int condition_host(kb_ctx* ctx, const double* X, long n, long hld, int p, const double* mu, const double* Sigma, const double* S,
                   int s_is_diag, int m, const double* Z, uint64_t seed, int64_t row_offset, double* Xko);

int gram_accumulate_host_rows(kb_ctx* ctx, const double* X, long n, long hld, int p, double* dG, long ld, double* dcs);

}  // namespace kb
