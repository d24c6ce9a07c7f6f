/*
 * CPU oracle (plain C) for the single-knockoff CCD s-solvers of Knockoffs.jl v2.0.2.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It restates, line by line, the reference's algorithm (paths relative to /root/reference):
 *   solve_MVR            src/utilities.jl:124-175
 *   forward_backward!    src/utilities.jl:182-185
 *   solve_quadratic      src/utilities.jl:187-199
 *   solve_max_entropy    src/utilities.jl:221-280
 *   solve_sdp_ccd        src/utilities.jl:291-343
 *   lowrankupdate_turbo! src/utilities.jl:516-563   (Givens, early termination)
 *   lowrankdowndate_turbo! src/utilities.jl:571-625
 * The reference is pure Julia and cannot run here; this file is pinned against the
 * reference's published known-answer vectors (tests/test_oracle_golden.py).
 *
 * Storage: the reference keeps the upper factor U (A = U'U) column-major, so its Givens row
 * loop A[i, i+1:n] is stride-p.  Here U is kept ROW-major (row i contiguous) so the same loop
 * vectorises like the reference's @turbo loop does; the arithmetic per element is identical.
 * Triangular solves (BLAS dtrsv in the reference) are written as plain loops; summation order
 * differs from OpenBLAS at rounding level only.
 *
 * Return codes: 0 ok, -1 invalid argument, -2 not positive definite (PosDefException),
 * -3 Inf/NaN delta (utilities.jl:195-196).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define U_(i, j) U[(size_t)(i) * (size_t)p + (size_t)(j)]

/* cholesky(Symmetric(A)) upper factor, row-major, in place on the upper triangle of U. */
static int chol_upper(double *U, int64_t p) {
    for (int64_t i = 0; i < p; ++i) {
        /* row i of U: U[i,j] = (A[i,j] - sum_{k<i} U[k,i] U[k,j]) / U[i,i] */
        for (int64_t k = 0; k < i; ++k) {
            const double uki = U_(k, i);
            if (uki == 0.0) continue;
            double *ri = &U_(i, 0);
            const double *rk = &U_(k, 0);
#pragma omp simd
            for (int64_t j = i; j < p; ++j) ri[j] -= uki * rk[j];
        }
        double d = U_(i, i);
        if (!(d > 0.0)) return -2;
        d = sqrt(d);
        U_(i, i) = d;
        const double inv = 1.0 / d;
        for (int64_t j = i + 1; j < p; ++j) U_(i, j) *= inv;
    }
    return 0;
}

/* x = U'^{-1} b (forward substitution).  `first` = index of the first nonzero of b. */
static void solve_Ut(const double *U, int64_t p, double *x, int64_t first) {
    /* x holds b on entry.  axpy form over contiguous rows of U. */
    for (int64_t k = first; k < p; ++k) {
        const double xk = x[k] / U_(k, k);
        x[k] = xk;
        if (xk == 0.0) continue;
        const double *rk = &U_(k, 0);
#pragma omp simd
        for (int64_t j = k + 1; j < p; ++j) x[j] -= xk * rk[j];
    }
}

/* x = U^{-1} y (backward substitution), dot form over contiguous rows. */
static void solve_U(const double *U, int64_t p, double *x) {
    for (int64_t i = p - 1; i >= 0; --i) {
        const double *ri = &U_(i, 0);
        double acc = 0.0;
#pragma omp simd reduction(+ : acc)
        for (int64_t j = i + 1; j < p; ++j) acc += ri[j] * x[j];
        x[i] = (x[i] - acc) / ri[i];
    }
}

/* utilities.jl:516-563; v = sqrt(|delta|) e_j, so idx_start = idx_end = j.  Destroys v. */
static void lowrankupdate_turbo(double *U, int64_t p, double *v, int64_t j, int robust,
                                double *touched) {
    int early = 0;
    for (int64_t i = j; i < p; ++i) {
        const double f = U_(i, i), g = v[i];
        double c, s, r;
        if (g == 0.0) { c = 1.0; s = 0.0; r = f; }
        else { r = hypot(f, g); c = f / r; s = g / r; }   /* givensAlgorithm, f > 0 */
        if (!robust) {
            if (fabs(s) < 1e-15 && i >= j) { if (++early > 10) break; } else early = 0;
        }
        U_(i, i) = r;
        double *ri = &U_(i, 0);
#pragma omp simd
        for (int64_t k = i + 1; k < p; ++k) {
            const double Aik = ri[k], vk = v[k];
            ri[k] = c * Aik + s * vk;
            v[k] = -s * Aik + c * vk;
        }
        *touched += (double)(p - i);
    }
}

/* utilities.jl:571-625 */
static int lowrankdowndate_turbo(double *U, int64_t p, double *v, int64_t j, int robust,
                                 double *touched) {
    int early = 0;
    for (int64_t i = j; i < p; ++i) {
        const double Aii = U_(i, i);
        const double s = v[i] / Aii;
        if (s * s > 1.0) return -2;                       /* PosDefException(i) */
        const double c = sqrt(1.0 - s * s);
        if (!robust) {
            if (fabs(s) < 1e-15 && i >= j) { if (++early > 10) break; } else early = 0;
        }
        U_(i, i) = c * Aii;
        double *ri = &U_(i, 0);
#pragma omp simd
        for (int64_t k = i + 1; k < p; ++k) {
            const double vk = v[k];
            const double Aik = (ri[k] - s * vk) / c;
            ri[k] = Aik;
            v[k] = -s * Aik + c * vk;
        }
        *touched += (double)(p - i);
    }
    return 0;
}

static int solve_quadratic(double cn, double cd, double Sjj, double m, double *out) {
    const double a = -cn - cd * cd * m * m;
    const double b = 2.0 * (-cn * Sjj + cd * m * m);
    const double c = -cn * Sjj * Sjj - m * m;
    if (a == 0.0 && c == 0.0) { *out = 0.0; return 0; }
    const double disc = sqrt(b * b - 4.0 * a * c);
    const double x1 = (-b + disc) / (2.0 * a);
    const double x2 = (-b - disc) / (2.0 * a);
    const double d = (-Sjj < x1 && x1 < 1.0 / cd) ? x1 : x2;
    if (isinf(d) || isnan(d)) return -3;
    *out = d;
    return 0;
}

static double sumsq(const double *x, int64_t from, int64_t p) {
    double acc = 0.0;
#pragma omp simd reduction(+ : acc)
    for (int64_t i = from; i < p; ++i) acc += x[i] * x[i];
    return acc;
}

/* Build row-major upper triangle of kappa*Sigma - diag(s) + shift*I from column-major Sigma
 * (only the upper triangle of Sigma is read: Symmetric(Sigma), uplo = :U). */
static double *build_factor(const double *Sigma, int64_t p, double kappa, const double *s,
                            double shift, int *rc) {
    double *U = (double *)malloc(sizeof(double) * (size_t)p * (size_t)p);
    if (!U) { *rc = -1; return NULL; }
    for (int64_t i = 0; i < p; ++i) {
        for (int64_t j = 0; j < i; ++j) U_(i, j) = 0.0;
        for (int64_t j = i; j < p; ++j) U_(i, j) = kappa * Sigma[(size_t)i + (size_t)j * p];
        U_(i, i) += shift - (s ? s[i] : 0.0);
    }
    *rc = chol_upper(U, p);
    if (*rc) { free(U); return NULL; }
    return U;
}

/* element (i,j) of Symmetric(Sigma) */
static inline double sym(const double *Sigma, int64_t p, int64_t i, int64_t j) {
    return i <= j ? Sigma[(size_t)i + (size_t)j * p] : Sigma[(size_t)j + (size_t)i * p];
}

/*
 * method: 0 = MVR, 1 = maxent, 2 = sdp_ccd.
 * Sigma: p x p column-major correlation matrix (upper triangle read).
 * s_init: required for MVR/ME (the caller evaluates solve_equi); ignored for sdp_ccd.
 * trace[niter]: per-sweep max|delta| (MVR/ME) or sum(s) (sdp_ccd); may be NULL.
 * stats[2]: [0] factor entries read+written by the rank-1 updates, [1] coordinates updated.
 * max_sweeps_override / max_coords_override (> 0): stop after that many sweeps / coordinates -- used
 * only by bench.py to time a BOUNDED sample of a large solve on the CPU.
 * U_init (optional): row-major upper factor of the initial matrix computed by the caller with LAPACK
 * dpotrf (what the reference's `cholesky` calls, utilities.jl:140); NULL = factor here.
 * stats[3]: [2] = seconds spent in the sweep loop (excludes the initial factorisation).
 * coord_stride (> 1): TIMING SAMPLE ONLY -- visit every coord_stride-th coordinate of the sweep.  A coordinate's cost
 * depends on its position j (the e_j solves and the rank-1 update touch rows j..p only), not on the state, so an evenly
 * spaced subset x coord_stride is an unbiased estimate of one whole sweep; the returned s is NOT a solution.
 */
int ko_solve_ccd(int method, const double *Sigma, int64_t p, double m, int niter, double tol,
                 double lambda_min, double sdp_lambda, double sdp_mu, int robust,
                 const double *s_init, double *s_out, int *sweeps_out, double *trace,
                 double *stats, int max_sweeps_override, int64_t max_coords_override,
                 const double *U_init, int64_t coord_stride) {
    if (p < 1 || m < 1.0 || niter < 0) return -1;
    if (method == 2 && (!(sdp_mu >= 0.0 && sdp_mu <= 1.0) || !(sdp_lambda > 0.0))) return -1;
    if (method != 2 && !s_init) return -1;
    const double kappa = (m + 1.0) / m;
    int rc = 0;
    double *s = s_out;
    if (method == 2) memset(s, 0, sizeof(double) * (size_t)p);
    else memcpy(s, s_init, sizeof(double) * (size_t)p);
    double *U;
    if (U_init) {
        U = (double *)malloc(sizeof(double) * (size_t)p * (size_t)p);
        if (!U) return -1;
        memcpy(U, U_init, sizeof(double) * (size_t)p * (size_t)p);
    } else {
        U = build_factor(Sigma, p, kappa, method == 2 ? NULL : s, method == 2 ? 0.0 : lambda_min, &rc);
        if (!U) return rc;
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    double *vd = (double *)calloc((size_t)p, sizeof(double));
    double *vn = (double *)calloc((size_t)p, sizeof(double));
    double *x = (double *)calloc((size_t)p, sizeof(double));
    double touched = 0.0, nupd = 0.0;
    double lam = sdp_lambda;
    int sweeps = 0, stop = 0;
    int64_t ncoord = 0;
    if (max_sweeps_override > 0 && max_sweeps_override < niter) niter = max_sweeps_override;
    for (int l = 0; l < niter && rc == 0; ++l) {
        ++sweeps;
        double max_delta = 0.0;
        for (int64_t j = 0; j < p && rc == 0; j += (coord_stride > 1 ? coord_stride : 1)) {
            if (max_coords_override > 0 && ncoord >= max_coords_override) { stop = 1; break; }
            ++ncoord;
            double delta;      /* signed change used for the factor: A <- A - delta e_j e_j' */
            if (method == 0) {
                memset(vd, 0, sizeof(double) * (size_t)p);
                vd[j] = 1.0;
                solve_Ut(U, p, vd, j);                          /* :152 */
                memcpy(vn, vd, sizeof(double) * (size_t)p);
                solve_U(U, p, vn);                              /* :149 */
                const double cn = -sumsq(vn, 0, p);
                const double cd = sumsq(vd, j, p);
                double dj;
                rc = solve_quadratic(cn, cd, s[j], m, &dj);     /* :155 */
                if (rc) break;
                const double ub = 1.0 / cd - lambda_min;        /* :157-158 */
                if (dj > ub) dj = ub;
                if (fabs(dj) < 1e-15) continue;                 /* :159 */
                s[j] += dj;
                delta = dj;
            } else {
                for (int64_t i = 0; i < p; ++i) x[i] = kappa * sym(Sigma, p, i, j);
                const double kSjj = x[j];
                x[j] = 0.0;
                solve_Ut(U, p, x, 0);                           /* :248 / :323 */
                const double x_l2 = sumsq(x, 0, p);
                const double zeta = kSjj - s[j];
                const double c = (zeta * x_l2) / (zeta + x_l2);
                if (method == 1) {
                    const double sj_new = (kSjj - c) / 2.0;     /* :254 */
                    memset(vd, 0, sizeof(double) * (size_t)p);
                    vd[j] = 1.0;
                    solve_Ut(U, p, vd, j);                      /* :258 */
                    const double ub = 1.0 / sumsq(vd, j, p) - lambda_min;
                    double d = sj_new - s[j];
                    if (d > ub) d = ub;
                    if (fabs(d) < 1e-15) continue;
                    s[j] = sj_new;                              /* :264 (unclipped, Q2) */
                    delta = d;
                } else {
                    double sj_new = kSjj - c - lam;             /* :329 */
                    sj_new = sj_new < 0.0 ? 0.0 : (sj_new > 1.0 ? 1.0 : sj_new);
                    const double d = s[j] - sj_new;             /* :330 */
                    if (fabs(d) < 1e-15) continue;
                    s[j] = sj_new;
                    delta = -d;                                 /* d>0 => update (A grows) */
                }
            }
            memset(vd, 0, sizeof(double) * (size_t)p);
            vd[j] = sqrt(fabs(delta));
            if (delta > 0.0) rc = lowrankdowndate_turbo(U, p, vd, j, robust, &touched);
            else lowrankupdate_turbo(U, p, vd, j, robust, &touched);
            nupd += 1.0;
            if (fabs(delta) > max_delta) max_delta = fabs(delta);
        }
        if (rc || stop) break;
        if (method == 2) {
            if (trace) { double t = 0; for (int64_t i = 0; i < p; ++i) t += s[i]; trace[l] = t; }
            lam *= sdp_mu;                                      /* :339-340 */
            if (lam < tol) break;
        } else {
            if (trace) trace[l] = max_delta;
            if (max_delta < tol) break;                          /* :172 / :277 */
        }
    }
    if (sweeps_out) *sweeps_out = sweeps;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (stats) {
        stats[0] = 2.0 * touched; stats[1] = nupd;
        stats[2] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    free(U); free(vd); free(vn); free(x);
    return rc;
}
