// Marginal feature-importance statistics + knockoff filter on the device.
// Reference: src/fit_lasso.jl:188-257, src/threshold.jl:19-31, 55-75, 96-142.
//
// The heavy part is one HBM-bound pass over X and X̃: T_c = abs2(zscore(A[:, c])' y_std) / n for every column c
// of [X  X̃] (8 n (m+1) p algorithmic bytes).  It is a ONE-pass, chunk-additive reduction: with the column's first
// entry c0 as a shift, s1 = Σ(x − c0), s2 = Σ(x − c0)², sy = Σ(x − c0)·ys give
//     mean = c0 + s1/n,   var = (s2 − s1²/n)/(n − 1),   X_std[:, c]'ys = (sy − (s1/n)·Σys)/sqrt(var),
// so row chunks can be accumulated while they are produced by the sampling GEMM and X̃ never has to be stored.
#include "stats.cuh"
#include <cmath>

namespace kb {

namespace {

constexpr int ST_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// deterministic CTA-wide sum (all threads get the result); red: >= 32 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

__global__ void __launch_bounds__(1024) zscore_vec_kernel(const double* __restrict__ y, long n, double* __restrict__ ys,
                                                           double* __restrict__ sum_out) {
    __shared__ double red[32];
    double a = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) a += y[i];
    const double mean = block_sum(a, red) / (double)n;
    a = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = y[i] - mean;
        a = fma(d, d, a);
    }
    const double sd = sqrt(block_sum(a, red) / (double)(n - 1));     // Julia's std: corrected
    a = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = (y[i] - mean) / sd;
        ys[i] = v;
        a += v;
    }
    const double s = block_sum(a, red);
    if (threadIdx.x == 0) sum_out[0] = s;
}

// one CTA per column
__global__ void __launch_bounds__(ST_THREADS) marginal_accumulate_kernel(const double* __restrict__ A, long rows, long lda,
                                                                          const double* __restrict__ ys,
                                                                          double* __restrict__ acc, int first) {
    __shared__ double red[32];
    const long c = blockIdx.x;
    const double* __restrict__ x = A + c * lda;
    const double c0 = first ? x[0] : acc[STATS_ACC * c];
    double s1 = 0.0, s2 = 0.0, sy = 0.0;
    long i = threadIdx.x;
    for (; i + 3 * ST_THREADS < rows; i += 4 * ST_THREADS) {
        const double x0 = x[i], x1 = x[i + ST_THREADS], x2 = x[i + 2 * ST_THREADS], x3 = x[i + 3 * ST_THREADS];
        const double y0 = ys[i], y1 = ys[i + ST_THREADS], y2 = ys[i + 2 * ST_THREADS], y3 = ys[i + 3 * ST_THREADS];
        const double d0 = x0 - c0, d1 = x1 - c0, d2 = x2 - c0, d3 = x3 - c0;
        s1 += (d0 + d1) + (d2 + d3);
        s2 = fma(d0, d0, s2); s2 = fma(d1, d1, s2); s2 = fma(d2, d2, s2); s2 = fma(d3, d3, s2);
        sy = fma(d0, y0, sy); sy = fma(d1, y1, sy); sy = fma(d2, y2, sy); sy = fma(d3, y3, sy);
    }
    for (; i < rows; i += ST_THREADS) {
        const double d = x[i] - c0;
        s1 += d;
        s2 = fma(d, d, s2);
        sy = fma(d, ys[i], sy);
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    sy = block_sum(sy, red);
    if (threadIdx.x == 0) {
        double* a = acc + STATS_ACC * c;
        if (first) { a[0] = c0; a[1] = s1; a[2] = s2; a[3] = sy; }
        else { a[1] += s1; a[2] += s2; a[3] += sy; }
    }
}

__global__ void marginal_finish_kernel(const double* __restrict__ acc, long ncols, double n, const double* __restrict__ sum_ys,
                                       double* __restrict__ T) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const double s1 = acc[STATS_ACC * c + 1], s2 = acc[STATS_ACC * c + 2], sy = acc[STATS_ACC * c + 3];
    const double mshift = s1 / n;                                  // mean − c0
    const double var = (s2 - s1 * mshift) / (n - 1.0);             // std(X, dims=1): corrected
    const double dot = (sy - mshift * sum_ys[0]) / sqrt(var);      // X_std[:, c]' y_std
    T[c] = dot * dot / n;                                          // abs2.(·) ./ n   (fit_lasso.jl:217)
}

__global__ void group_abs_sum_kernel(const double* __restrict__ T, long p, int nblocks, const int* __restrict__ gidx,
                                     int ng, double* __restrict__ Tg) {
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long)nblocks * ng) return;
    const int g = (int)(id % ng);
    const long k = id / ng;
    const double* t = T + k * p;
    double a = 0.0;
    for (long j = 0; j < p; ++j)
        if (gidx[j] == g) a += fabs(t[j]);
    Tg[id] = a;
}

constexpr int MK_MAX_M = 64;

// median of m values sorted in DESCENDING order (Statistics.median: middle(a, b) = a/2 + b/2 for even m)
__device__ __forceinline__ double median_sorted(const double* v, int m) {
    if (m & 1) return v[m / 2];
    return v[m / 2 - 1] / 2 + v[m / 2] / 2;
}

__global__ void mk_statistics_kernel(const double* __restrict__ T0, const double* __restrict__ Tk, long p, int m,
                                     long long* __restrict__ kappa, double* __restrict__ tau, double* __restrict__ W) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p) return;
    const double t0 = fabs(T0[i]);
    if (m == 1) {                                                  // threshold.jl:134-142
        W[i] = t0 - fabs(Tk[i]);
        return;
    }
    double st[MK_MAX_M + 1];
    long long kap = 0;
    for (int k = 0; k < m; ++k) {
        const double tk = fabs(Tk[(long)k * p + i]);
        if (tk > t0) kap = k + 1;                                  // the LAST such k (threshold.jl:108-110)
        // insertion into st[0..k) kept descending
        int pos = k;
        while (pos > 0 && st[pos - 1] < tk) { st[pos] = st[pos - 1]; --pos; }
        st[pos] = tk;
    }
    const double med_ko = median_sorted(st, m);
    W[i] = (t0 - med_ko) * (kap == 0 ? 1.0 : 0.0);                 // threshold.jl:113
    // sort!(storage, rev=true): insert |T0| into the descending list, then largest − median(rest)
    int pos = m;
    while (pos > 0 && st[pos - 1] < t0) { st[pos] = st[pos - 1]; --pos; }
    st[pos] = t0;
    tau[i] = st[0] - median_sorted(st + 1, m);                     // threshold.jl:114-115
    kappa[i] = kap;
}

__global__ void set_inf_kernel(double* out) { out[0] = __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ void atomic_min_pos(double* addr, double v) {   // v > 0
    atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void threshold_kernel(const double* __restrict__ w, long p, double q, int offset, long rej_bounds,
                                 double* __restrict__ out) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double t = fabs(w[j]);
    if (!(t > 0.0)) return;
    long neg = 0, pos = 0, greater = 0;
    for (long k = 0; k < p; ++k) {
        const double x = w[k];
        neg += (x <= -t);
        pos += (x >= t);
        greater += (fabs(x) > t);
    }
    if (greater > rej_bounds) return;                              // `i > rej_bounds && break` (threshold.jl:29)
    const double ratio = (double)(offset + neg) / (double)pos;     // threshold.jl:27
    if (ratio <= q) atomic_min_pos(out, t);
}

__global__ void mk_threshold_kernel(const double* __restrict__ tau, const long long* __restrict__ kappa, long p, int m,
                                    double q, long rej_bounds, double* __restrict__ out) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double t = tau[j];
    if (!(t > 0.0)) return;                                        // `0 < t` (threshold.jl:71)
    long numer = 0, denom = 0, greater = 0;
    for (long k = 0; k < p; ++k) {
        const double x = tau[k];
        const bool ge = (x >= t);
        numer += (ge && kappa[k] >= 1);
        denom += (ge && kappa[k] == 0);
        greater += (x > t);
    }
    if (greater > rej_bounds) return;
    const double offset = 1.0 / (double)m;
    // (offset + offset * numer_counter) / max(1, demon_counter)   (threshold.jl:70) -- no FMA contraction
    const double ratio = __ddiv_rn(__dadd_rn(offset, __dmul_rn(offset, (double)numer)), (double)(denom > 1 ? denom : 1));
    if (ratio <= q) atomic_min_pos(out, t);
}

__global__ void select_kernel(const double* __restrict__ W, long p, const double* __restrict__ tau_hat, int* __restrict__ sel) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) sel[j] = (W[j] >= tau_hat[0]) ? 1 : 0;
}

}  // namespace

int zscore_vec_dev(kb_ctx* ctx, const double* dy, long n, double* dys, double* d_sum_ys) {
    if (n < 2) return set_error(ctx, KB_ERR_INVALID, "zscore: need at least 2 observations");
    zscore_vec_kernel<<<1, 1024, 0, ctx->stream>>>(dy, n, dys, d_sum_ys);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int marginal_accumulate_dev(kb_ctx* ctx, const double* dA, long rows, long lda, long ncols, const double* dys, double* dacc,
                            int first) {
    if (rows <= 0 || ncols <= 0) return KB_OK;
    if (ncols > 0x7fffffffL) return set_error(ctx, KB_ERR_INVALID, "marginal statistics: too many columns");
    marginal_accumulate_kernel<<<(unsigned)ncols, ST_THREADS, 0, ctx->stream>>>(dA, rows, lda, dys, dacc, first);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int marginal_finish_dev(kb_ctx* ctx, const double* dacc, long ncols, long n, const double* d_sum_ys, double* dT) {
    marginal_finish_kernel<<<ceil_div(ncols, 256), 256, 0, ctx->stream>>>(dacc, ncols, (double)n, d_sum_ys, dT);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int group_abs_sum_dev(kb_ctx* ctx, const double* dT, long p, int nblocks, const int* dgidx, int ng, double* dTg) {
    group_abs_sum_kernel<<<ceil_div((long)nblocks * ng, 128), 128, 0, ctx->stream>>>(dT, p, nblocks, dgidx, ng, dTg);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int mk_statistics_dev(kb_ctx* ctx, const double* dT0, const double* dTk, long p, int m, long long* dkappa, double* dtau,
                      double* dW) {
    if (m < 1 || m > MK_MAX_M) return set_error(ctx, KB_ERR_INVALID, "MK_statistics: m must be in 1..%d but was %d", MK_MAX_M, m);
    if (p <= 0) return KB_OK;
    mk_statistics_kernel<<<ceil_div(p, 128), 128, 0, ctx->stream>>>(dT0, dTk, p, m, dkappa, dtau, dW);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int threshold_dev(kb_ctx* ctx, const double* dw, long p, double q, int offset, long rej_bounds, double* d_out) {
    if (!(q >= 0.0 && q <= 1.0))                                   // threshold.jl:21
        return set_error(ctx, KB_ERR_INVALID, "Target FDR should be between 0 and 1 but got %g", q);
    set_inf_kernel<<<1, 1, 0, ctx->stream>>>(d_out);
    ctx->launches++;
    if (p > 0) {
        threshold_kernel<<<ceil_div(p, 128), 128, 0, ctx->stream>>>(dw, p, q, offset, rej_bounds, d_out);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int mk_threshold_dev(kb_ctx* ctx, const double* dtau, const long long* dkappa, long p, int m, double q, long rej_bounds,
                     double* d_out) {
    if (!(q >= 0.0 && q <= 1.0))                                   // threshold.jl:58
        return set_error(ctx, KB_ERR_INVALID, "Target FDR should be between 0 and 1 but got %g", q);
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    set_inf_kernel<<<1, 1, 0, ctx->stream>>>(d_out);
    ctx->launches++;
    if (p > 0) {
        mk_threshold_kernel<<<ceil_div(p, 128), 128, 0, ctx->stream>>>(dtau, dkappa, p, m, q, rej_bounds, d_out);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int select_dev(kb_ctx* ctx, const double* dW, long p, const double* d_tau_hat, int* dsel) {
    if (p <= 0) return KB_OK;
    select_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(dW, p, d_tau_hat, dsel);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace kb
