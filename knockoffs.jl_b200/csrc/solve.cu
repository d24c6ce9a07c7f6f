// solve_s on the device: cov->cor, s_init (solve_equi), initial factorisation, CCD kernel, rescale.
// Reference: src/utilities.jl:27-51 (solve_s), :104-111 (solve_equi), :124-343 (solvers).
#include "kb_common.cuh"
#include "linalg.cuh"
#include "ccd.cuh"
#include "pipeline.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace kb {

__global__ void extract_sigma_kernel(const double* Sig, long lds, int p, double* sig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p) sig[i] = sqrt(Sig[i + (long)i * lds]);
}

// Cor (lower, ld) from the UPPER triangle of Sig (Symmetric(Σ), uplo = :U).
// iscor: copy as is.  Otherwise StatsBase.cov2cor semantics (utilities.jl:33): off-diagonals
// Σ_ik / (σ_i σ_k) clamped to [-1, 1], unit diagonal.
__global__ void build_cor_lower_kernel(const double* Sig, long lds, int p, const double* sig, int iscor,
                                       double* Cor, long ld) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bk = blockIdx.y;   // output tile rows bi, cols bk  (bi >= bk)
    if (bi < bk) return;
    // read the transposed tile: element (k, i) of Sig, k in tile bk, i in tile bi
    const int kr = bk * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bi * 32 + r;
        t[r][threadIdx.x] = (kr < p && i < p && kr <= i) ? Sig[kr + (long)i * lds] : 0.0;
    }
    __syncthreads();
    const int i = bi * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = bk * 32 + r;
        if (i < p && k < p && i >= k) {
            double v = t[threadIdx.x][r];
            if (!iscor) {
                if (i == k) v = 1.0;
                else {
                    v = v / (sig[i] * sig[k]);
                    v = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);
                }
            }
            Cor[i + (long)k * ld] = v;
        }
    }
}

// A (lower) = kappa * Cor with diagonal kappa*Cor_jj - s_j + shift (init != 0: also records kappa*Cor_jj and A_jj)
// or, for init == 0, with the tracked diagonal adiag (refresh of the inverse between sweeps).
__global__ void build_A_lower_kernel(const double* Cor, long ld, int p, double kappa, const double* s, double shift,
                                     double* A, double* ksig_diag, double* adiag, int init) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)ld * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % ld), k = (int)(idx / ld);
        if (i < k || i >= p) { A[i + (long)k * ld] = 0.0; continue; }
        double v = kappa * Cor[i + (long)k * ld];
        if (i == k) {
            if (init) {
                if (ksig_diag) ksig_diag[i] = v;
                v = v - (s ? s[i] : 0.0) + shift;
                adiag[i] = v;
            } else {
                v = adiag[i];
            }
        }
        A[i + (long)k * ld] = v;
    }
}

__global__ void fill_kernel(double* x, int n, double v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void scale_s_kernel(double* s, const double* sig, int p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p) s[i] *= sig[i] * sig[i];
}
// Gershgorin lower bound and min diagonal of a symmetric matrix stored in its lower triangle
__global__ void gershgorin_kernel(const double* B, long ld, int p, double* out /*[0]=min lb,[1]=min diag*/) {
    __shared__ double s_lb[256], s_d[256];
    double lb = INFINITY, dmin = INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p; i += gridDim.x * blockDim.x) {
        double off = 0.0;
        for (int k = 0; k < i; ++k) off += fabs(B[i + (long)k * ld]);
        for (int k = i + 1; k < p; ++k) off += fabs(B[k + (long)i * ld]);
        const double d = B[i + (long)i * ld];
        lb = fmin(lb, d - off);
        dmin = fmin(dmin, d);
    }
    s_lb[threadIdx.x] = lb; s_d[threadIdx.x] = dmin;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_lb[threadIdx.x] = fmin(s_lb[threadIdx.x], s_lb[threadIdx.x + o]);
            s_d[threadIdx.x] = fmin(s_d[threadIdx.x], s_d[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = s_lb[0]; out[2 * blockIdx.x + 1] = s_d[0]; }
}
__global__ void shift_copy_lower_kernel(const double* B, long ld, int p, double sigma, double* A) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        A[i + (long)k * ld] = (i < k) ? 0.0 : (B[i + (long)k * ld] - (i == k ? sigma : 0.0));
    }
}


// ---- inverse iteration on a successful probe's factor: Rayleigh-quotient UPPER bounds for λmin ----------------------
// t_part[ks][i] = scale * Σ_{k in chunk ks, k <= i} W[i,k] x[k]      (W lower triangular, column-major)
__global__ void tri_gemv_lower_kernel(const double* __restrict__ W, long ldw, int p, const double* __restrict__ x, double scale,
                                      int kchunk, double* __restrict__ part, long ldp) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int k0 = blockIdx.y * kchunk;
    if (i >= p) return;
    const int k1 = min(min(p, k0 + kchunk), i + 1);
    double a0 = 0.0, a1 = 0.0;
    int k = k0;
    for (; k + 1 < k1; k += 2) {
        a0 = fma(W[i + (long)k * ldw], x[k], a0);
        a1 = fma(W[i + (long)(k + 1) * ldw], x[k + 1], a1);
    }
    if (k < k1) a0 = fma(W[i + (long)k * ldw], x[k], a0);
    part[(long)blockIdx.y * ldp + i] = (a0 + a1) * scale;
}
__global__ void reduce_parts_kernel(const double* __restrict__ part, long ldp, int ks, int p, double* __restrict__ t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p) return;
    double a = 0.0;
    for (int k = 0; k < ks; ++k) a += part[(long)k * ldp + i];
    t[i] = a;
}
// y_k = Σ_{i >= k} W[i,k] t_i : one warp per column
__global__ void tri_gemv_lower_T_kernel(const double* __restrict__ W, long ldw, int p, const double* __restrict__ t,
                                        double* __restrict__ y) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= p) return;
    const double* w = W + (long)k * ldw;
    double a = 0.0;
    for (int i = k + lane; i < p; i += 32) a = fma(w[i], t[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) y[k] = a;
}
// colq[k] = y_k (B_kk y_k + 2 Σ_{i > k} B_ik y_i)  (B symmetric, lower triangle stored);  Σ_k colq[k] = y'By
__global__ void quadform_lower_kernel(const double* __restrict__ B, long ld, int p, const double* __restrict__ y,
                                      double* __restrict__ colq) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= p) return;
    const double* b = B + (long)k * ld;
    double a = 0.0;
    for (int i = k + 1 + lane; i < p; i += 32) a = fma(b[i], y[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) colq[k] = y[k] * (b[k] * y[k] + 2.0 * a);
}

// Reductions behind the objective values (deterministic: fixed per-block partials, summed on the host).
// mode 0: part[b] = Σ x_i² over `count` contiguous doubles;  mode 1: part[b] = Σ_j log(x[j * stride]) (j < count);
// mode 2: part[b] = Σ_j log(x_j + shift);  mode 3: part[b] = Σ_j 1 / x_j;  mode 4: part[b] = Σ_j x_j.
__global__ void objective_reduce_kernel(const double* __restrict__ x, long count, long stride, double shift, int mode,
                                        double* __restrict__ part) {
    __shared__ double red[256];
    double a = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        if (mode == 0) { const double v = x[i]; a = fma(v, v, a); }
        else if (mode == 1) a += log(x[i * stride]);
        else if (mode == 2) a += log(x[i] + shift);
        else if (mode == 3) a += 1.0 / x[i];
        else a += x[i];
    }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

static inline int ewg(kb_ctx* ctx, long total) {
    long g = (total + 255) / 256, cap = (long)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Root λ of the power-law model h(σ) = C (λ − σ)^α through three points σ0 < σ1 < σ2 with h0 > h1 > h2 > 0
// (α = 1: an isolated smallest eigenvalue, the last pivot is a simple pole's reciprocal; α = 1/2: the edge of a
// near-continuous spectrum such as AR(1)).  Returns NaN when the points do not fit the model.
static double powerlaw_root(const double x[3], const double y[3]) {
    if (!(y[0] > y[1] && y[1] > y[2] && y[2] > 0.0)) return NAN;
    const double R = (log(y[0]) - log(y[1])) / (log(y[1]) - log(y[2]));
    auto F = [&](double lam) { return (log(lam - x[0]) - log(lam - x[1])) / (log(lam - x[1]) - log(lam - x[2])) - R; };
    double a = x[2] + (x[2] - x[1]) * 1e-12, b = x[2] + (x[2] - x[0]) * 1e6;
    double fa = F(a);
    const double fb = F(b);
    if (!(fa == fa) || !(fb == fb) || fa * fb > 0.0) return NAN;
    for (int it = 0; it < 200; ++it) {
        const double mid = x[2] + sqrt((a - x[2]) * (b - x[2]));     // bisection on the log of the distance
        const double fm = F(mid);
        if (!(fm == fm)) return NAN;
        if (fa * fm <= 0.0) b = mid; else { a = mid; fa = fm; }
        if (b - a <= 1e-15 * fabs(b)) break;
    }
    return 0.5 * (a + b);
}

// λmin of the symmetric matrix B (lower triangle, ld): B − σI ≻ 0  <=>  σ < λmin, tested with the blocked DMMA
// Cholesky, so every probe is rigorous; the bracket [lo (PD), hi (not PD)] is closed to 2e-13 relative.
// Probes are placed by extrapolating the LAST PIVOT h(σ) = L_pp(σ)² = 1/[(B − σI)⁻¹]_pp of the successful
// factorisations (concave, decreasing, zero at λmin): a power-law / secant root estimate is probed just above
// (expect failure -> hi) or just below (expect success -> lo) by the change of the estimate; any probe that fails
// to halve the bracket is followed by a plain bisection step, so the worst case is ~bisection (≈ 45
// factorisations) and isolated / multiple smallest eigenvalues take ≈ 13.  No host LAPACK anywhere.
int eigmin_lower_dev(kb_ctx* ctx, const double* B, long ld, int p, double* work_A, double* invdiag, int* d_fail,
                     double* lambda_out, bool inverse_iteration) {
    double* d_g = nullptr;
    Scope sc(ctx);
    KB_TRY(sc.alloc(&d_g, 2 * 64));
    gershgorin_kernel<<<64, 256, 0, ctx->stream>>>(B, ld, p, d_g);
    ctx->launches++;
    double hg[128];
    KB_CUDA(ctx, cudaMemcpyAsync(hg, d_g, sizeof(hg), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double lb = INFINITY, dmin = INFINITY;
    for (int i = 0; i < 64; ++i) { lb = fmin(lb, hg[2 * i]); dmin = fmin(dmin, hg[2 * i + 1]); }
    if (!std::isfinite(lb) || !std::isfinite(dmin)) return set_error(ctx, KB_ERR_NAN, "eigmin: non-finite matrix");
    // probe: pd = (B − σI ≻ 0); h = last pivot (valid when pd)
    auto is_pd = [&](double sigma, bool* pd, double* h) -> int {
        shift_copy_lower_kernel<<<ewg(ctx, (long)p * p), 256, 0, ctx->stream>>>(B, ld, p, sigma, work_A);
        ctx->launches++;
        KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
        KB_TRY(potrf_lower(ctx, work_A, ld, p, invdiag, d_fail));
        int f = 0;
        double lpp = 0.0;
        KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaMemcpyAsync(&lpp, work_A + (size_t)(p - 1) + (size_t)(p - 1) * ld, sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *pd = (f == 0);
        *h = lpp * lpp;
        return KB_OK;
    };
    const double tol = 2e-13;
    static const bool trace_probes = getenv("KB_EIGMIN_TRACE") != nullptr;
    double lo, hi = dmin;   // λmin <= min diagonal
    double px[3] = {0, 0, 0}, py[3] = {0, 0, 0};   // last PD probes, ascending σ
    int npts = 0;
    auto push = [&](double x, double y) {
        if (npts == 3) { px[0] = px[1]; py[0] = py[1]; px[1] = px[2]; py[1] = py[2]; npts = 2; }
        px[npts] = x; py[npts] = y; ++npts;
    };
    bool pd = false;
    double h = 0.0;
    if (lb > 0.0) {
        lo = lb;                                    // Gershgorin: B − lb I is PSD; probe it for h
        KB_TRY(is_pd(lo, &pd, &h));
        if (pd) push(lo, h); else lo = lb - 1e-12 * fabs(lb) - 1e-300;
    } else {
        KB_TRY(is_pd(0.0, &pd, &h));
        if (pd) { lo = 0.0; push(0.0, h); }
        else { lo = lb - 1e-12 * fabs(lb) - 1e-300; hi = fmin(hi, 0.0); }
    }
    if (hi <= lo) { *lambda_out = hi; return KB_OK; }
    // Inverse iteration on the factor of the first successful probe: x <- (B − lo I)⁻¹ x through W = L⁻¹ (two
    // triangular matrix-vector products per step, HBM-bound); every Rayleigh quotient ρ(x) = x'Bx / x'x is an UPPER
    // bound of λmin.  With a spectral gap above λmin (any multiplicity: block-equicorrelated Σ) it converges in a
    // handful of steps and ONE more probe just below it closes the bracket (p = 5000: 11 -> 3 factorisations);
    // without a gap (AR(1) edge) it still replaces the first bisection steps.  The probes stay the only arbiter.
    static const bool use_rq = [] { const char* e = getenv("KB_EIGMIN_RQ"); return !(e && e[0] == '0'); }();
    const bool rq_on = inverse_iteration && use_rq && p >= 2;
    const int KS = 16;
    const int kchunk = (p + KS - 1) / KS;
    double *Wi = nullptr, *wk = nullptr, *part = nullptr, *xv = nullptr, *tv = nullptr, *yv = nullptr, *colq = nullptr;
    std::vector<double> hy, hq;
    const double* xin = nullptr;
    double xscale = 1.0, stage_width = INFINITY;
    if (rq_on) {
        KB_TRY(sc.alloc(&Wi, (size_t)ld * p));
        KB_TRY(sc.alloc(&wk, (size_t)ld * p / 2 + 2 * ld + 64));
        KB_TRY(sc.alloc(&part, (size_t)KS * ld));
        KB_TRY(sc.alloc(&xv, (size_t)ld));
        KB_TRY(sc.alloc(&tv, (size_t)ld));
        KB_TRY(sc.alloc(&yv, (size_t)ld));
        KB_TRY(sc.alloc(&colq, (size_t)ld));
        hy.resize(p); hq.resize(p);
        std::vector<double> hx(p);
        for (int i = 0; i < p; ++i) hx[i] = 1.0 + 0.37 * sin(1.0 + 2.39996322972865332 * i);   // fixed, generic start
        KB_CUDA(ctx, cudaMemcpyAsync(xv, hx.data(), sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        xin = xv;
    }
    // One stage: work_A must hold the factor of B − lo I (the last probe was the successful one at lo).
    auto rq_stage = [&]() -> int {
        stage_width = hi - lo;
        KB_TRY(trtri_lower(ctx, work_A, ld, p, invdiag, Wi, ld, wk));
        double rho_prev = NAN, d_prev = NAN, rho = NAN, est_err = NAN, qr_last = 1.0;
        int its = 0;
        for (; its < 40; ++its) {
            dim3 g1(ceil_div(p, 128), KS);
            tri_gemv_lower_kernel<<<g1, 128, 0, ctx->stream>>>(Wi, ld, p, xin, xscale, kchunk, part, ld);
            reduce_parts_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(part, ld, KS, p, tv);
            tri_gemv_lower_T_kernel<<<ceil_div(p, 8), 256, 0, ctx->stream>>>(Wi, ld, p, tv, yv);
            quadform_lower_kernel<<<ceil_div(p, 8), 256, 0, ctx->stream>>>(B, ld, p, yv, colq);
            ctx->launches += 4;
            KB_CUDA(ctx, cudaMemcpyAsync(hy.data(), yv, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaMemcpyAsync(hq.data(), colq, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            long double yy = 0.0L, q = 0.0L;
            for (int i = 0; i < p; ++i) { yy += (long double)hy[i] * hy[i]; q += hq[i]; }
            if (!(yy > 0.0L) || !std::isfinite((double)yy) || !std::isfinite((double)q)) return KB_OK;
            rho = (double)(q / yy);
            xscale = 1.0 / sqrt((double)yy);
            xin = yv;   // next step reads y (scaled inside the kernel); y is rewritten only after t has been formed
            const double dcur = rho_prev - rho;                      // >= 0 up to rounding
            if (its >= 1) {
                double qr = (d_prev == d_prev && d_prev > 0.0) ? dcur / d_prev : NAN;   // ≈ r², r = (λmin − lo)/(λ2 − lo)
                if (!(qr == qr) || qr < 0.0) qr = 0.0;
                qr_last = qr;
                if (qr > 0.95) qr = 0.95;
                est_err = fabs(dcur) * qr / (1.0 - qr);
                if (fabs(dcur) <= 4e-16 * fabs(rho)) { est_err = 0.0; ++its; break; }
                if (its >= 8 && qr > 0.8) { ++its; break; }           // no gap at this shift: leave the rest to the probes
            }
            d_prev = dcur; rho_prev = rho;
        }
        if (trace_probes) fprintf(stderr, "[eigmin rq] %d steps rho=%.17g est_err=%.3e\n", its, rho, est_err);
        if (rho == rho && rho > lo && rho * (1.0 + 64.0 * 2.220446049250313e-16) < hi)
            hi = rho * (1.0 + 64.0 * 2.220446049250313e-16) + 1e-300;   // upper bound, padded for the rounding of x'Bx
        if (!(rho == rho) || !(rho > lo)) return KB_OK;
        // two probes below the bound: a safe one (the Rayleigh quotient of an eigenvector of a matrix with condition
        // number c carries ~c·eps of rounding), then one at a quarter of the final tolerance
        const double margins[2] = {fmax(4.0 * (est_err == est_err ? est_err : 0.0), 2.5e-12 * fabs(rho)), 0.25 * tol * fabs(rho)};
        if (margins[0] > 0.5 * (hi - lo) || (est_err != 0.0 && qr_last > 0.5)) return KB_OK;   // slow / unconverged: the error
                                                                      // estimate is not trustworthy, keep only the bound
        for (int t = 0; t < 2; ++t) {
            const double sigma = rho - margins[t];
            if (!(sigma > lo && sigma < hi)) break;
            KB_TRY(is_pd(sigma, &pd, &h));
            if (pd) { lo = sigma; if (t == 0) npts = 0; push(sigma, h); } else hi = sigma;
            if (trace_probes)
                fprintf(stderr, "[eigmin rq] probe sigma=%.17g %s width=%.3e\n", sigma, pd ? "PD" : "no", hi - lo);
            if (!pd || margins[0] > 1e-9 * fabs(rho)) break;          // not converged: hand over to the bracket loop
        }
        return KB_OK;
    };
    // Inverse iteration on the factor of a successful probe: x <- (B − lo I)⁻¹ x through W = L⁻¹ (two triangular
    // matrix-vector products per step, HBM-bound); every Rayleigh quotient ρ(x) = x'Bx / x'x is an UPPER bound of λmin.
    // With a spectral gap above λmin (any multiplicity: block-equicorrelated Σ) the first stage converges in a handful
    // of steps and two more probes just below ρ close the bracket (p = 5000: 11 -> 3 factorisations).  Without a gap
    // (AR(1) edge) the stage is repeated whenever the bracket has shrunk 100x -- it converges once lo is within the
    // gap of λmin (~22 instead of 46 factorisations).  The Cholesky probes stay the only arbiter.
    if (rq_on && pd) KB_TRY(rq_stage());
    double est_prev = NAN;
    bool force_bisect = false;
    for (int it = 0; it < 200; ++it) {
        if (hi - lo <= tol * fmax(fabs(lo), fabs(hi))) break;
        const double w0 = hi - lo;
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        double est = NAN;
        if (npts >= 2) {
            const double a = px[npts - 2], ha = py[npts - 2], b = px[npts - 1], hb = py[npts - 1];
            if (ha > hb && hb > 0.0) est = b + hb * (b - a) / (ha - hb);       // secant root (>= λmin for concave h)
        }
        if (npts == 3 && est == est) {
            // the curvature-corrected root must lie between the last success and the secant root; anything else is
            // an ill-conditioned fit (three nearly collinear points) and is discarded
            const double pl = powerlaw_root(px, py);
            if (pl == pl && pl > px[2] && pl <= est) est = pl;
        }
        double sigma = mid;
        bool guided = false;
        if (est == est && est >= hi && est - hi <= 0.01 * w0) est = hi;   // root modelled at a known failure: aim just below it
        if (est == est && est > lo && est <= hi && !force_bisect) {
            double unc = (est_prev == est_prev) ? fabs(est - est_prev) : 0.25 * (est - lo);
            unc = fmax(unc, 0.25 * tol * fabs(est));
            sigma = (hi - est > 4.0 * unc) ? est + 2.0 * unc : est - 2.0 * unc;
            guided = (sigma > lo && sigma < hi);
            if (!guided) sigma = mid;
        }
        if (est == est) est_prev = est;
        KB_TRY(is_pd(sigma, &pd, &h));
        if (pd) { lo = sigma; push(sigma, h); } else hi = sigma;
        force_bisect = guided && (hi - lo) > 0.5 * w0;
        if (rq_on && pd && (hi - lo) <= 0.01 * stage_width && (hi - lo) > tol * fmax(fabs(lo), fabs(hi))) KB_TRY(rq_stage());
        if (trace_probes)
            fprintf(stderr, "[eigmin %2d] %s sigma=%.17g %s h=%.6e est=%.17g width=%.3e\n", it, guided ? "guided" : "bisect", sigma,
                    pd ? "PD" : "no", pd ? h : 0.0, est, hi - lo);
    }
    *lambda_out = 0.5 * (lo + hi);
    return KB_OK;
}

int eigmin_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, double* lambda_out) {
    const long ld = (p + 1) & ~1L;
    Scope sc(ctx);
    double *B, *A, *invdiag, *d_one;
    int* d_fail;
    KB_TRY(sc.alloc(&B, (size_t)ld * p));
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&d_one, p));
    KB_TRY(sc.alloc(&d_fail, 4));
    dim3 grid(ceil_div(p, 32), ceil_div(p, 32)), block(32, 8);
    build_cor_lower_kernel<<<grid, block, 0, ctx->stream>>>(dSigma, lds, p, d_one, 1, B, ld);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return eigmin_lower_dev(ctx, B, ld, p, A, invdiag, d_fail, lambda_out, true);
}

void default_solve_opts(kb_solve_opts* o) {
    o->niter = 100; o->tol = 1e-6; o->lambda_min = 1e-6; o->sdp_lambda = 0.5; o->sdp_mu = 0.8;
    o->robust = 0; o->verbose = 0; o->mode = KB_MODE_AUTO; o->refresh_every = -1; o->objective = 1;
}

// dSigma: p x p, leading dimension lds, upper triangle trusted.  d_s_init may be NULL.  d_s_out: p.
//
// Host loop = the reference's `for l in 1:niter` (utilities.jl:143, :240, :311): one persistent-kernel
// launch per sweep; between sweeps the host reads max|δ| (the reference's convergence test) and, unless
// opts.refresh_every == 0, rebuilds M = A⁻¹ from the current diagonal of A on the tensor-core
// Cholesky/inverse (≈ p³ flops, a few % of a sweep) so Sherman–Morrison rounding never accumulates
// over sweeps.  refresh_every == 0 runs the whole solve inside ONE launch.
int solve_s_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, int method, double m, const kb_solve_opts* opts_in,
                const double* d_s_init, double* d_s_out, kb_solve_info* info) {
    kb_solve_opts opts;
    if (opts_in) opts = *opts_in; else default_solve_opts(&opts);
    if (info) { memset(info, 0, sizeof(*info)); }
    if (p < 1) return set_error(ctx, KB_ERR_INVALID, "p must be >= 1 but was %d", p);
    if (!(m >= 1.0)) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %g.", m);   // utilities.jl:28
    if (method != KB_METHOD_MVR && method != KB_METHOD_MAXENT && method != KB_METHOD_SDP_CCD && method != KB_METHOD_EQUI)
        return set_error(ctx, KB_ERR_INVALID, "Method must be one of mvr/maxent/sdp_ccd/equi but was %d", method);
    if (method == KB_METHOD_SDP_CCD) {
        if (!(opts.sdp_mu >= 0.0 && opts.sdp_mu <= 1.0))
            return set_error(ctx, KB_ERR_INVALID, "Decay parameter mu must be in [0, 1] but was %g", opts.sdp_mu);
        if (!(opts.sdp_lambda > 0.0))
            return set_error(ctx, KB_ERR_INVALID, "Barrier coefficient lambda must be > 0 but was %g", opts.sdp_lambda);
    }
    if (opts.niter < 0) return set_error(ctx, KB_ERR_INVALID, "niter must be >= 0");
    const unsigned long long launches0 = ctx->launches;
    const double flops0 = ctx->flops;
    const double kappa = (m + 1.0) / m;
    const long ld = ccd_stream_ld(p);
    Scope sc(ctx);
    double *d_sig, *Cor, *A, *W, *Mbuf, *work, *invdiag, *ksig, *adiag, *s, *colbuf, *trace, *counters;
    int *d_fail, *d_status, *d_sweeps;
    unsigned long long* d_ready;
    KB_TRY(sc.alloc(&d_sig, p));
    KB_TRY(sc.alloc(&Cor, (size_t)ld * p));
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&ksig, p));
    KB_TRY(sc.alloc(&adiag, p));
    KB_TRY(sc.alloc(&s, p));
    KB_TRY(sc.alloc(&trace, KB_MAX_TRACE));
    KB_TRY(sc.alloc(&counters, 8));
    KB_TRY(sc.alloc(&d_fail, 4));
    KB_TRY(sc.alloc(&d_status, 4));
    KB_TRY(sc.alloc(&d_sweeps, 4));
    KB_TRY(sc.alloc(&d_ready, 4));

    // ---- cov -> cor (utilities.jl:31-33) ----
    extract_sigma_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(dSigma, lds, p, d_sig);
    ctx->launches++;
    std::vector<double> hsig(p);
    KB_CUDA(ctx, cudaMemcpyAsync(hsig.data(), d_sig, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool iscor = true;
    for (int i = 0; i < p; ++i) {
        const double x = hsig[i];
        if (!(x == x) || x <= 0.0) return set_error(ctx, KB_ERR_NOT_PD, "Sigma has a non-positive diagonal entry at %d", i + 1);
        // Julia isapprox(x, 1): |x - 1| <= sqrt(eps) * max(|x|, 1)
        if (!(fabs(x - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(x), 1.0))) iscor = false;
    }
    {
        dim3 grid(ceil_div(p, 32), ceil_div(p, 32)), block(32, 8);
        build_cor_lower_kernel<<<grid, block, 0, ctx->stream>>>(dSigma, lds, p, d_sig, iscor ? 1 : 0, Cor, ld);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
    }

    // ---- s_init ----
    const bool need_equi = (method == KB_METHOD_EQUI) || (method != KB_METHOD_SDP_CCD && d_s_init == nullptr);
    if (need_equi) {
        double lam = 0.0;
        KB_TRY(eigmin_lower_dev(ctx, Cor, ld, p, A, invdiag, d_fail, &lam, true));
        double sj = kappa * lam;                           // utilities.jl:108-110
        if (sj > 1.0) sj = 1.0;
        fill_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(s, p, sj);
        ctx->launches++;
    } else if (method == KB_METHOD_SDP_CCD) {
        KB_TRY(fill_zero(ctx, s, p));                      // utilities.jl:308
    } else {
        KB_CUDA(ctx, cudaMemcpyAsync(s, d_s_init, sizeof(double) * p, cudaMemcpyDeviceToDevice, ctx->stream));
    }

    int status = KB_OK;
    if (method != KB_METHOD_EQUI && opts.niter > 0) {
        KB_TRY(sc.alloc(&W, (size_t)ld * p));
        KB_TRY(sc.alloc(&Mbuf, (size_t)ld * p));
        KB_TRY(sc.alloc(&work, (size_t)ld * p / 2 + 2 * ld + 64));
        // execution mode: AUTO = the blocked (delayed rank-64 update) kernels; KB_MODE_STREAM = the persistent
        // rank-1 streaming kernel.  KB_CCD_MODE=stream|block overrides AUTO (experiments).
        int mode = (opts.mode == KB_MODE_STREAM) ? KB_MODE_STREAM : KB_MODE_BLOCK;
        if (opts.mode == KB_MODE_AUTO) {
            const char* e = getenv("KB_CCD_MODE");
            if (e && e[0] == 's') mode = KB_MODE_STREAM;
        }
        double* blockwork = nullptr;
        if (mode == KB_MODE_STREAM) KB_TRY(sc.alloc(&colbuf, ccd_stream_colbuf_doubles(p)));
        else KB_TRY(sc.alloc(&blockwork, ccd_block_work_doubles(p)));
        KB_TRY(fill_zero(ctx, trace, KB_MAX_TRACE));
        KB_TRY(fill_zero(ctx, counters, 8));
        const double shift = (method == KB_METHOD_SDP_CCD) ? 0.0 : opts.lambda_min;
        const double pd = (double)p;
        // rebuild cadence of the resident inverse.  Default: every sweep for the streaming kernel (p rank-1
        // passes per sweep); for the blocked one (p/64 rank-64 passes per sweep) after the FIRST sweep -- the equi
        // starting point makes the initial A nearly singular (λmin(A) = λmin option), the inverse built there is the
        // least accurate one of the whole solve -- and then every 8th sweep (measured at p = 5000: same objective to 1e-15
        // as every 4th, and runs with no refresh at all still meet the 1e-8 parity tests).  :sdp_ccd starts from s = 0
        // with no shift, i.e. from a well conditioned A = κΣ: its 59-sweep solve at p = 5000 ends 1.76e-12 from the
        // golden s whether the inverse is rebuilt every 8 sweeps or never (profiles/r02_refresh_sweep.json), so the
        // blocked kernel rebuilds it every 32nd sweep only (8.5 ms per rebuild).
        const bool sdp = (method == KB_METHOD_SDP_CCD);
        const bool auto_block = (opts.refresh_every < 0 && mode == KB_MODE_BLOCK && !sdp);
        const int refresh_every = opts.refresh_every < 0 ? (mode == KB_MODE_STREAM ? 1 : (sdp ? 32 : 8)) : opts.refresh_every;
        double lam = opts.sdp_lambda;
        double init_ms = 0.0, solve_ms = 0.0;
        int sweeps_done = 0;
        std::vector<double> hs(p);
        bool converged = false;
        double block_touched = 0.0;

        // (re)build M = A^{-1}; first == true also initialises ksig / adiag from s (utilities.jl:140, :237, :309)
        auto build_inverse = [&](bool first, bool* ok) -> int {
            KB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
            build_A_lower_kernel<<<ewg(ctx, (long)ld * p), 256, 0, ctx->stream>>>(Cor, ld, p, kappa, first ? s : nullptr, shift,
                                                                               A, first ? ksig : nullptr, adiag, first ? 1 : 0);
            ctx->launches++;
            KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
            KB_TRY(potrf_lower(ctx, A, ld, p, invdiag, d_fail));
            int f = 0;
            KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            *ok = (f == 0);
            if (f != 0) {
                if (first)
                    return set_error(ctx, KB_ERR_NOT_PD, "PosDefException: matrix is not positive definite; "
                                                         "Cholesky factorization failed at pivot %d.", f);
                return KB_OK;   // keep the incrementally updated inverse
            }
            KB_TRY(trtri_lower(ctx, A, ld, p, invdiag, W, ld, work));
            KB_TRY(fill_zero(ctx, Mbuf, (size_t)ld * p));
            KB_TRY(syrk_WtW_lower(ctx, W, ld, p, Mbuf, ld));
            KB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            KB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            init_ms += ms;
            return KB_OK;
        };

        bool ok = false;
        KB_TRY(build_inverse(true, &ok));

        CcdParams P{};
        P.p = p; P.method = method; P.m = m; P.tol = opts.tol; P.lambda_min = opts.lambda_min;
        P.sdp_mu = opts.sdp_mu;
        P.M = Mbuf; P.ld = ld; P.s = s; P.adiag = adiag; P.ksig_diag = ksig; P.colbuf = colbuf; P.ready = d_ready;
        P.status = d_status; P.sweeps_out = d_sweeps; P.trace = trace; P.counters = counters;

        while (sweeps_done < opts.niter && !converged && status == KB_OK) {
            const int chunk = (mode == KB_MODE_BLOCK) ? 1
                              : (refresh_every == 0) ? (opts.niter - sweeps_done)
                                                     : std::min(refresh_every, opts.niter - sweeps_done);
            if (sweeps_done > 0 && refresh_every > 0 && (sweeps_done % refresh_every == 0 || (auto_block && sweeps_done == 1)))
                KB_TRY(build_inverse(false, &ok));
            P.nsweeps = chunk; P.sweep0 = sweeps_done; P.sdp_lambda = lam;
            KB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
            if (mode == KB_MODE_STREAM) {
                KB_TRY(ccd_stream_run(ctx, P, nullptr));
            } else {
                KB_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(int), ctx->stream));
                KB_CUDA(ctx, cudaMemsetAsync(d_sweeps, 0, sizeof(int), ctx->stream));
                KB_TRY(ccd_block_run(ctx, P, blockwork));
                // bytes the rank-64 passes move: every block reads + writes the lower 128 x 128 tiles once
                const double nt = (double)((p + 127) / 128);
                block_touched += (double)((p + CCD_BLOCK_WIDTH - 1) / CCD_BLOCK_WIDTH) * (nt * (nt + 1) / 2) * 128.0 * 128.0 * 16.0;
            }
            KB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
            int hsweeps = 0;
            double hcount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double htrace[KB_MAX_TRACE];
            KB_CUDA(ctx, cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaMemcpyAsync(&hsweeps, d_sweeps, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaMemcpyAsync(hcount, counters, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaMemcpyAsync(htrace, trace, sizeof(double) * KB_MAX_TRACE, cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaMemcpyAsync(hs.data(), s, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
            KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            KB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            solve_ms += ms;
            if (status == KB_ERR_CUDA) {
                set_error(ctx, status, "CCD kernel watchdog: a CTA waited > 2 s for a column (aborted).");
                break;
            }
            const int last = sweeps_done + hsweeps - 1;
            sweeps_done += hsweeps;
            lam = hcount[2];
            if (info) {
                // MVR / maxent: the device trace holds max|δ| per sweep; sdp_ccd: sum(s) per sweep is filled in below
                if (method != KB_METHOD_SDP_CCD)
                    for (int l = 0; l < sweeps_done && l < KB_MAX_TRACE; ++l) info->trace[l] = htrace[l];
                info->updates = hcount[1];
                info->touched_bytes = (mode == KB_MODE_STREAM) ? 16.0 * CCD_SLOT_ROWS * hcount[0] : block_touched;
                if (getenv("KB_CCD_PHASES"))
                    fprintf(stderr, "[ccd launch %d: %.2f ms; phases, Mcycles summed over CTAs so far] wait=%.1f stage=%.1f "
                                    "publish=%.1f stream=%.1f slots=%.0f\n", sweeps_done, ms,
                            hcount[4] / 1e6, hcount[5] / 1e6, hcount[6] / 1e6, hcount[7] / 1e6, hcount[0]);
            }
            if (status != KB_OK) break;
            if (method == KB_METHOD_SDP_CCD) {
                if (info && last >= 0 && last < KB_MAX_TRACE) {   // what verbose prints: sum(s) (utilities.jl:313-316)
                    double t = 0.0;
                    for (int i = 0; i < p; ++i) t += hs[i];
                    // a launch of the streaming kernel may cover several sweeps: only its last sweep's s is visible here
                    for (int l = last - hsweeps + 1; l <= last; ++l)
                        if (l >= 0) info->trace[l] = t;
                }
                if (lam < opts.tol) converged = true;              // utilities.jl:339-340
            } else if (hcount[3] < opts.tol) {
                converged = true;                                  // utilities.jl:172 / :277
            }
            if (hsweeps < chunk) converged = true;                 // the kernel itself broke out of its sweep loop
        }
        if (info) {
            info->init_ms = init_ms;
            info->solve_ms = solve_ms;
            info->sweeps = sweeps_done;
            info->mode = mode;
            // full-sweep algorithmic bytes of the reference's solves + update
            const double per_sweep = (method == KB_METHOD_SDP_CCD) ? (20.0 / 3.0) * pd * pd * pd : 8.0 * pd * (pd + 1) * (pd + 1);
            info->ref_bytes = per_sweep * sweeps_done;
        }
        if (status == KB_ERR_NOT_PD) set_error(ctx, status, "PosDefException: rank-1 downdate left the positive definite cone.");
        else if (status == KB_ERR_NAN) set_error(ctx, status, "delta_j is Inf/NaN, aborting.");
        // ---- objective of the returned s on the correlation scale (north-star "MVR trace / ME logdet / SDP sum"; the
        //      single-knockoff solvers of the reference never evaluate one, so the definitions are group_block_objective's,
        //      group.jl:24 (ME, with its 1e-8 shifts) and :27 (MVR), and sum(s) for sdp_ccd, utilities.jl:79).
        //      Evaluated from a FRESH factorisation of κΣ − S (not from the Sherman–Morrison-updated inverse, which
        //      also carries the λmin shift): one Cholesky (ME) plus one triangular inverse (MVR).  NaN if κΣ − S is not PD.
        if (info && status == KB_OK && opts.objective) {
            double* part;
            KB_TRY(sc.alloc(&part, 256));
            auto reduce = [&](const double* x, long count, long stride, double shift_, int rmode, double* out) -> int {
                objective_reduce_kernel<<<256, 256, 0, ctx->stream>>>(x, count, stride, shift_, rmode, part);
                ctx->launches++;
                double h[256];
                KB_CUDA(ctx, cudaMemcpyAsync(h, part, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
                KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                long double t = 0.0L;
                for (int i = 0; i < 256; ++i) t += h[i];
                *out = (double)t;
                return KB_OK;
            };
            double obj = NAN;
            if (method == KB_METHOD_SDP_CCD) {
                KB_TRY(reduce(s, p, 1, 0.0, 4, &obj));
            } else {
                const double oshift = (method == KB_METHOD_MAXENT) ? 1e-8 : 0.0;
                build_A_lower_kernel<<<ewg(ctx, (long)ld * p), 256, 0, ctx->stream>>>(Cor, ld, p, kappa, s, oshift, A, nullptr, adiag, 1);
                ctx->launches++;
                KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
                KB_TRY(potrf_lower(ctx, A, ld, p, invdiag, d_fail));
                int f = 0;
                KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (f == 0) {
                    double t1 = 0.0, t2 = 0.0;
                    if (method == KB_METHOD_MAXENT) {
                        KB_TRY(reduce(A, p, ld + 1, 0.0, 1, &t1));          // Σ log L_jj
                        KB_TRY(reduce(s, p, 1, 1e-8, 2, &t2));              // Σ log(s_j + 1e-8)
                        obj = 2.0 * t1 + m * t2;
                    } else {
                        KB_TRY(trtri_lower(ctx, A, ld, p, invdiag, W, ld, work));
                        KB_TRY(reduce(W, (long)ld * p, 1, 0.0, 0, &t1));    // ‖L⁻¹‖_F² = tr((κΣ − S)⁻¹)
                        KB_TRY(reduce(s, p, 1, 0.0, 3, &t2));               // Σ 1 / s_j
                        obj = m * m * t2 + t1;
                    }
                }
            }
            info->objective = obj;
        }
    }
    // ---- rescale (utilities.jl:49) ----
    if (!iscor) {
        scale_s_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(s, d_sig, p);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaMemcpyAsync(d_s_out, s, sizeof(double) * p, cudaMemcpyDeviceToDevice, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info) {
        info->status = status;
        info->launches = (int)(ctx->launches - launches0);
        info->flops = ctx->flops - flops0;
    }
    return status;
}

}  // namespace kb
