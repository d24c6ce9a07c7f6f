// approx_modelX_gaussian_knockoffs on the device: the covariance is approximated by a block
// diagonal matrix of windows of adjacent features; every window is an independent (Ledoit-Wolf Σ̂_w, solve_s)
// problem and `condition` factorises window by window.  Reference: src/approx.jl:37-109.
//
// Two phases so that windows can be dealt across GPUs (they are embarrassingly parallel):
//   approx_solve_windows      Σ̂_w = cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X[:, w]), s_w = solve_s(Σ̂_w),
//                             γ_w = sup{γ in [0,1] : κΣ̂_w − γ·Diag(s_w) ≻ 0}                 (approx.jl:80-97, 105-109)
//   approx_condition_windows  X̃[:, w] = condition(X[:, w], μ_w, Σ̂_w, Diag(γ s_w); m)       (approx.jl:99-100; Σ, Σ⁻¹S, C
//                             and every factor block of modelX.jl:106-120 are block diagonal, so the p x p
//                             algebra of the reference is exactly the union of the windows' algebra)
#include "approx.cuh"
#include "linalg.cuh"
#include "pipeline.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace kb {

namespace {

// A (lower, strict upper zero, ld lda) = κ·Symmetric(Σ) − γ·Diag(s);  Σ full symmetric with ld lds
__global__ void gamma_probe_kernel(const double* __restrict__ Sig, long lds, int p, double kappa, double gamma,
                                   const double* __restrict__ s, double* __restrict__ A, long lda) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        double v = 0.0;
        if (i >= k) {
            v = kappa * Sig[k + (long)i * lds];                    // upper triangle is the trusted one
            if (i == k) v -= gamma * s[i];
        }
        A[i + (long)k * lda] = v;
    }
}
__global__ void scale_kernel(double* x, long n, double a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= a;
}

int ew_grid(kb_ctx* ctx, long total) {
    long g = (total + 255) / 256, cap = (long)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// sup{γ in [0,1] : κΣ − γ·Diag(s) ≻ 0} for one dense Σ (device, upper triangle trusted).
// `fzero(γ -> f(γ, s, Σ, m), 0, 1)` of approx.jl:97 with f = 1 − γ where λmin(κΣ − γD) > 0, −Inf elsewhere:
// γ = 1 if that is feasible, else the bisection limit (returned from the feasible side).  The positive-definiteness
// probe is a Cholesky factorisation instead of `eigmin`.
int feasible_gamma_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* ds, double m, double* gamma_out) {
    const long ld = (p + 1) & ~1L;
    const double kappa = (m + 1.0) / m;
    Scope sc(ctx);
    double *A, *invdiag;
    int* d_fail;
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&d_fail, 4));
    auto feasible = [&](double gamma, bool* ok) -> int {
        gamma_probe_kernel<<<ew_grid(ctx, (long)p * p), 256, 0, ctx->stream>>>(dSigma, lds, p, kappa, gamma, ds, A, ld);
        ctx->launches++;
        KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
        KB_TRY(potrf_lower(ctx, A, ld, p, invdiag, d_fail));
        int f = 0;
        KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *ok = (f == 0);
        return KB_OK;
    };
    bool ok = false;
    KB_TRY(feasible(1.0, &ok));
    if (ok) { *gamma_out = 1.0; return KB_OK; }
    KB_TRY(feasible(0.0, &ok));
    if (!ok) return set_error(ctx, KB_ERR_NOT_PD, "approx: the window covariance is not positive definite");
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;                        // adjacent floating-point numbers
        KB_TRY(feasible(mid, &ok));
        if (ok) lo = mid; else hi = mid;
    }
    *gamma_out = lo;
    return KB_OK;
}

// Phase 1 for windows [w0, w1).  dX: the columns win_off[w0] .. win_off[w1] of X on the device (n rows, ld ldx).  win_off: nwin + 1 ascending 0-based column
// offsets (window w = columns win_off[w] .. win_off[w+1]).  Outputs (device): dSigblk = the windows' Σ̂_w packed one
// after the other (window w: pw x pw column-major, ld pw, at offset Σ_{v<w} pv²), ds / dmu (p; only the entries of
// the handled windows are written).  Host: gamma_min_out = min over the handled windows of γ_w.
int approx_solve_windows(kb_ctx* ctx, const double* dX, long n, long ldx, long p, const int64_t* win_off, int nwin, int w0,
                         int w1, int method, double m, const kb_solve_opts* opts, double* dSigblk, double* ds, double* dmu,
                         double* gamma_min_out, kb_solve_info* info) {
    if (info) memset(info, 0, sizeof(*info));
    long pmax = 0;
    std::vector<size_t> boff((size_t)nwin + 1, 0);
    for (int w = 0; w < nwin; ++w) {
        const long pw = (long)(win_off[w + 1] - win_off[w]);
        pmax = std::max(pmax, pw);
        boff[(size_t)w + 1] = boff[(size_t)w] + (size_t)pw * pw;
    }
    const long ldm = (pmax + 1) & ~1L;
    Scope sc(ctx);
    double *dG, *dG2, *dY;
    KB_TRY(sc.alloc(&dG, (size_t)ldm * pmax));
    KB_TRY(sc.alloc(&dG2, (size_t)ldm * pmax));
    const long ychunk = std::min(n, 32768L);
    const long ldy = (ychunk + 1) & ~1L;
    KB_TRY(sc.alloc(&dY, (size_t)ldy * pmax));
    double gmin = 1.0;
    for (int w = w0; w < w1; ++w) {
        const long a = (long)win_off[w];
        const int pw = (int)(win_off[w + 1] - win_off[w]);
        const long ld = (pw + 1) & ~1L;
        const double* Xw = dX + (size_t)(a - (long)win_off[w0]) * ldx;
        double* dmean = dmu + a;
        // Σ̂_w = cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X[:, w])              (approx.jl:84)
        KB_TRY(fill_zero(ctx, dG, (size_t)ld * pw));
        KB_TRY(fill_zero(ctx, dG2, (size_t)ld * pw));
        KB_TRY(fill_zero(ctx, dmean, (size_t)pw));
        for (long r0 = 0; r0 < n; r0 += ychunk)
            KB_TRY(gram_partial_dev(ctx, Xw + r0, std::min(ychunk, n - r0), ldx, pw, dG, ld, dmean, 1));
        KB_TRY(gram_finish_dev(ctx, dG, ld, dmean, n, pw, 1, (double)n));                    // dmean: sums -> means
        for (long r0 = 0; r0 < n; r0 += ychunk)
            KB_TRY(gram_sq_partial_dev(ctx, Xw + r0, std::min(ychunk, n - r0), ldx, pw, dmean, dY, ldy, dG2, ld, 1));
        double lam = 0.0;
        KB_TRY(lw_finish_dev(ctx, dG, ld, dG2, ld, pw, n, &lam));
        double* Sw = dSigblk + boff[(size_t)w];
        KB_TRY(copy_matrix(ctx, dG, ld, Sw, pw, pw, pw));
        // s_w = solve_s(Σ̂_w, method; m, kwargs...)                                         (approx.jl:86)
        kb_solve_info wi;
        KB_TRY(solve_s_dev(ctx, Sw, (long)pw, pw, method, m, opts, nullptr, ds + a, &wi));
        if (info) {
            info->sweeps = std::max(info->sweeps, wi.sweeps);
            info->mode = wi.mode;
            info->solve_ms += wi.solve_ms; info->init_ms += wi.init_ms;
            info->ref_bytes += wi.ref_bytes; info->touched_bytes += wi.touched_bytes;
            info->flops += wi.flops; info->updates += wi.updates; info->objective += wi.objective;
            info->launches += wi.launches;
        }
        // γ: bisection search over [0, 1] to ensure Diag(γ s) ≤ κΣ                          (approx.jl:96-97, 105-109);
        // Σ is block diagonal, so the global γ is the minimum of the windows' γ_w
        double gw = 1.0;
        KB_TRY(feasible_gamma_dev(ctx, Sw, (long)pw, pw, ds + a, m, &gw));
        gmin = std::min(gmin, gw);
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *gamma_min_out = gmin;
    return KB_OK;
}

// Phase 2 for windows [w0, w1) (dX as in phase 1): ds already holds γ·s.  X̃ columns of window w, copy k go to host Xko (n x m p,
// ld n) at column k p + win_off[w]; Z (host, n x m p, same layout) may be NULL (device Philox stream keyed by the
// GLOBAL (row, column) of `randn(n, m p)`, so the result does not depend on how windows are dealt to GPUs).
int approx_condition_windows(kb_ctx* ctx, const double* dX, long n, long ldx, long p, const int64_t* win_off, int nwin,
                             int w0, int w1, int m, const double* dSigblk, const double* ds, const double* dmu,
                             const double* Z, uint64_t seed, int64_t row_offset, double* Xko) {
    long pmax = 0;
    std::vector<size_t> boff((size_t)nwin + 1, 0);
    for (int w = 0; w < nwin; ++w) {
        const long pw = (long)(win_off[w + 1] - win_off[w]);
        pmax = std::max(pmax, pw);
        boff[(size_t)w + 1] = boff[(size_t)w] + (size_t)pw * pw;
    }
    Scope sc(ctx);
    const long ldn = (n + 1) & ~1L;
    double *dSiS, *dLb, *dZw, *dOut;
    KB_TRY(sc.alloc(&dSiS, (size_t)pmax * pmax));
    KB_TRY(sc.alloc(&dLb, (size_t)pmax * pmax * (2 * m - 1)));
    KB_TRY(sc.alloc(&dZw, (size_t)ldn * pmax * m));
    KB_TRY(sc.alloc(&dOut, (size_t)ldn * pmax * m));
    for (int w = w0; w < w1; ++w) {
        const long a = (long)win_off[w];
        const int pw = (int)(win_off[w + 1] - win_off[w]);
        KB_TRY(cond_factor_dev(ctx, dSigblk + boff[(size_t)w], (long)pw, pw, ds + a, (long)pw, 1, m, dSiS, dLb, (long)pw));
        for (int k = 0; k < m; ++k) {
            double* Zk = dZw + (size_t)k * pw * ldn;
            if (Z) {
                KB_CUDA(ctx, cudaMemcpy2DAsync(Zk, ldn * sizeof(double), Z + (size_t)((long)k * p + a) * n, n * sizeof(double),
                                               n * sizeof(double), pw, cudaMemcpyHostToDevice, ctx->stream));
            } else {
                KB_TRY(randn_dev(ctx, seed, row_offset, n, pw, (long)k * p + a, Zk, ldn));
            }
        }
        KB_TRY(sample_dev(ctx, dX + (size_t)(a - (long)win_off[w0]) * ldx, n, ldx, pw, dmu + a, dSiS, dLb, (long)pw, m, dZw, ldn, seed, row_offset,
                          dOut, ldn));
        for (int k = 0; k < m; ++k)
            KB_CUDA(ctx, cudaMemcpy2DAsync(Xko + (size_t)((long)k * p + a) * n, n * sizeof(double), dOut + (size_t)k * pw * ldn,
                                           ldn * sizeof(double), n * sizeof(double), pw, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int scale_vec_dev(kb_ctx* ctx, double* dx, long n, double a) {
    if (n <= 0) return KB_OK;
    scale_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(dx, n, a);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace kb
