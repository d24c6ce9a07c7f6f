/*
 * libknockoffs_b200 -- C ABI of the B200-native (sm_100a) model-X Gaussian knockoff hot path.
 *
 * Drop-in boundary for biona001/Knockoffs.jl v2.0.2 (paths below are relative to that repo):
 * Julia keeps argument handling and struct construction; everything from `solve_s(...)`
 * (src/modelX.jl:62) through `condition(...)` (src/modelX.jl:64) is one of the calls below.
 * INTEGRATION.md shows the `ccall` stubs a maintainer would add.
 *
 * Conventions
 *   - all matrices are column-major FP64 with leading dimension = number of rows
 *     (Julia `Matrix{Float64}`); sizes are int64_t; `groups` are 1-based int64 ids.
 *   - Sigma is read from its UPPER triangle only (`Symmetric(Sigma)`, default uplo = :U) and
 *     no input is ever written (reference guarantee: test/runtests.jl:582-583).
 *   - host entry points take HOST pointers, copy in/out and hold no pointer after return;
 *     `_dev` variants take DEVICE pointers on the context's device and run on ctx's stream.
 *   - every function returns an int status (below); kb_last_error() gives the message.
 *     Nothing aborts, nothing prints unless opts->verbose.
 *   - one kb_ctx per host thread; distinct contexts may be used concurrently.
 */
#ifndef KNOCKOFFS_B200_H
#define KNOCKOFFS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (map to the reference's exceptions) ------------------------------------ */
#define KB_OK 0
#define KB_ERR_INVALID (-1) /* error("m should be 1 or larger ...") modelX.jl:105, utilities.jl:28,46,301-302 */
#define KB_ERR_NOT_PD (-2)  /* PosDefException: initial cholesky utilities.jl:140,237,309; downdate :591-593 */
#define KB_ERR_NAN (-3)     /* "δj is Inf/NaN, aborting" utilities.jl:195-196 */
#define KB_ERR_CUDA (-4)    /* CUDA / NCCL / out-of-memory */

/* ---- methods ----------------------------------------------------------------------------- */
#define KB_METHOD_MVR 0     /* :mvr     -> solve_MVR          utilities.jl:124-175 */
#define KB_METHOD_MAXENT 1  /* :maxent  -> solve_max_entropy  utilities.jl:221-280 */
#define KB_METHOD_SDP_CCD 2 /* :sdp_ccd -> solve_sdp_ccd      utilities.jl:291-343 (group :sdp likewise is CCD) */
#define KB_METHOD_EQUI 3    /* :equi    -> solve_equi         utilities.jl:104-111 */

/* solver execution modes (same mathematics, different kernels) */
#define KB_MODE_AUTO 0
#define KB_MODE_STREAM 1  /* persistent kernel, one rank-1 streaming pass over the resident triangle per coordinate */
#define KB_MODE_BLOCK 2   /* delayed updates: 64 coordinates per rank-64 tensor-core pass over the triangle (AUTO picks this) */

#define KB_MAX_TRACE 128

typedef struct kb_ctx kb_ctx;

/* Options of solve_s's keyword arguments (utilities.jl:124-133, 221-230, 291-300).
 * kb_default_solve_opts() fills the reference defaults. */
typedef struct {
    int32_t niter;       /* 100 */
    double tol;          /* 1e-6 */
    double lambda_min;   /* λmin = 1e-6 (MVR / maxent) */
    double sdp_lambda;   /* λ = 0.5 (sdp_ccd barrier) */
    double sdp_mu;       /* μ = 0.8 (sdp_ccd decay) */
    int32_t robust;      /* accepted, no effect on results beyond rounding */
    int32_t verbose;     /* per-sweep trace is always returned in kb_solve_info */
    int32_t mode;        /* KB_MODE_* */
    int32_t refresh_every; /* rebuild the device-resident inverse from the current diagonal every k sweeps
                              (-1 = default: 1 in KB_MODE_STREAM; in KB_MODE_BLOCK after the first sweep, then every 8th; 0 = never) */
    int32_t objective;   /* 1 (default): evaluate kb_solve_info.objective at exit (one extra factorisation of κΣ − S);
                            0: skip it */
} kb_solve_opts;

typedef struct {
    int32_t sweeps;               /* sweeps executed */
    int32_t status;               /* same as the return code */
    int32_t mode;                 /* mode that actually ran */
    int32_t launches;             /* kernels launched by the call */
    double trace[KB_MAX_TRACE];   /* per sweep: max|δ| (MVR, maxent) or sum(s) (sdp_ccd) -- what verbose prints */
    double solve_ms;              /* device time of the CCD kernel(s), CUDA events */
    double init_ms;               /* device time of the initial factorisation */
    double ref_bytes;             /* algorithmic bytes of the reference's steps actually taken */
    double touched_bytes;         /* bytes this implementation's update passes read + wrote */
    double flops;                 /* FP64 flops the call issued to the tensor-core (DMMA) GEMM kernel: inverse builds,
                                     eigmin probes, blocked sweeps, objective */
    double updates;               /* coordinates whose |δ| >= 1e-15 (factor updates performed) */
    double objective;             /* objective at the returned s, on the correlation scale Σcor the solver works in, with the
                                     definitions of group_block_objective (group.jl:24,27) / utilities.jl:79:
                                     :mvr m²·Σ1/s_j + tr((κΣcor − S)⁻¹); :maxent logdet(κΣcor − S + 1e-8 I) + m·Σlog(s_j + 1e-8);
                                     :sdp_ccd Σ s_j.  NaN when κΣcor − S is not positive definite; 0 when opts.objective == 0 */
} kb_solve_info;

/* Keyword arguments of solve_group_{max_entropy,mvr,sdp}_hybrid (group.jl:870-881, 951-962, 1057-1068). */
typedef struct {
    int32_t outer_iter;      /* 100 */
    int32_t inner_pca_iter;  /* 1 */
    int32_t inner_ccd_iter;  /* 1 */
    double tol;              /* 1e-4: |Δobj/obj| < tol or max δ < 1e-4 ends an inner loop */
    double eps;              /* ϵ = 1e-6 added to the feasible-interval bounds */
    int32_t robust;          /* accepted, no effect */
    int32_t verbose;
} kb_group_opts;

typedef struct {
    int32_t outer_iters;              /* outer iterations executed */
    int32_t inner_sweeps;             /* PCA + CCD sweeps executed (entries of the traces) */
    int32_t n_directions;             /* columns of the PCA direction set V */
    int32_t status;
    double obj_init, obj;             /* accumulated objective as the reference tracks it (obj_new += change_obj) */
    double gamma_equi;                /* γ of solve_group_equi used by initialize_S */
    double trace_obj[KB_MAX_TRACE];   /* per inner sweep: what `verbose` prints (group.jl:1198-1206, ...) */
    double trace_delta[KB_MAX_TRACE];
    int32_t trace_kind[KB_MAX_TRACE]; /* 0 = PCA sweep, 1 = CCD sweep */
    double steps, accepted;           /* 1-D line searches done / accepted */
    double solve_ms;                  /* device time of the sweeps (CUDA events) */
    double touched_bytes;             /* bytes the rank-1/2 streaming updates read + wrote */
} kb_group_info;

/* ---- context ----------------------------------------------------------------------------- */
int kb_create(kb_ctx** ctx, int device);
/* A context for THROUGHPUT work that is to run beside another context's latency-bound solve on the same GPU (e.g. sampling
 * the :mvr knockoffs while the :sdp_ccd solve of configs[2] is sweeping): all of its kernels -- main stream included -- are
 * confined to the SM partition that leaves KB_GREEN_SMS (default 16) SMs free and run at the lowest stream priority, so the
 * single-CTA chain kernels of the foreground context always find a free SM.  Falls back to an ordinary low-priority context
 * when the driver has no green contexts.  Same API as a kb_create context. */
int kb_create_background(kb_ctx** ctx, int device);
void kb_destroy(kb_ctx* ctx);
const char* kb_last_error(const kb_ctx* ctx);
const char* kb_version(void);
void kb_default_solve_opts(kb_solve_opts* opts);
/* kernels launched by this library on ctx since creation (bench.py's gpu_launches) */
uint64_t kb_launch_count(const kb_ctx* ctx);
/* the cudaStream_t all of ctx's kernels are launched on (for CUDA-event timing by the caller) */
void* kb_stream(const kb_ctx* ctx);
/* SMs the context's auxiliary stream is confined to (CUDA green context: the latency-bound single-CTA chain kernels of
 * the main stream always find a free SM beside the throughput work); 0 = no partition.  KB_GREEN_SMS=<n> sets the number of
 * SMs kept free (default 16, 0 disables). */
int32_t kb_aux_sms(const kb_ctx* ctx);
int kb_synchronize(kb_ctx* ctx);
/* release the device workspace the context keeps between calls */
int kb_trim(kb_ctx* ctx);
/* `condition` factors with cholesky(PositiveFactorizations.Positive, ·) (modelX.jl:111, :120), which never throws: a
 * non-positive pivot is replaced instead.  That package is not vendored with the reference, so its algorithm is RESTATED
 * from its published description (csrc/positive.cu; parity unpinned) and is opt-in: on != 0 makes the conditional-factor
 * stage fall back to it when the ordinary Cholesky fails (e.g. :equi, whose κΣ − S is exactly singular); the default
 * (off) keeps returning KB_ERR_NOT_PD.  For positive definite input both give the ordinary factor.
 * kb_positive_modified: pivots replaced by the last kb_cond_factor / kb_condition / one-shot call on ctx (0 = none). */
int kb_set_positive_cholesky(kb_ctx* ctx, int32_t on);
int32_t kb_positive_modified(const kb_ctx* ctx);
/* rows per chunk of the sampling pipeline (0 = default 4096) */
int kb_set_sample_chunk_rows(kb_ctx* ctx, int64_t rows);

/* ---- (1) s-solvers ------------------------------------------------------------------------
 * kb_solve_s replaces `solve_s(Σ::Symmetric, method; m, kwargs...)` (utilities.jl:27-51):
 * cov->cor, dispatch to MVR / maxent / sdp_ccd / equi, rescale by σ².  `s_init` (length p, on the
 * correlation scale) replaces the `s_init` kwarg; NULL = `solve_equi` (utilities.jl:130,227).
 */
int kb_solve_s(kb_ctx* ctx, const double* Sigma, int64_t p, int32_t method, double m,
               const kb_solve_opts* opts, const double* s_init, double* s_out, kb_solve_info* info);
int kb_solve_s_dev(kb_ctx* ctx, const double* dSigma, int64_t p, int32_t method, double m,
                   const kb_solve_opts* opts, const double* d_s_init, double* d_s_out, kb_solve_info* info);

/* λmin(Symmetric(Sigma)) -- the `eigvals(Σ) |> minimum` of solve_equi (utilities.jl:108). */
int kb_eigmin(kb_ctx* ctx, const double* Sigma, int64_t p, double* lambda_min_out);

/* ---- (1b) group s-solvers ------------------------------------------------------------------
 * kb_solve_s_group replaces `solve_s_group(Σ::Symmetric, groups, method; m, kwargs...)` (group.jl:348-422)
 * for method = KB_METHOD_MAXENT / KB_METHOD_MVR / KB_METHOD_SDP_CCD (group `:sdp` IS coordinate descent,
 * group.jl:403-404) / KB_METHOD_EQUI.  groups: p 1-based ids in any order (sorted contiguous inside, S is
 * returned in the caller's order).  V (p x nv, column-major, in the caller's variable order, each column
 * supported on one group) replaces get_PCA_vectors' direction set; NULL = computed here
 * (per-group eigenvectors, block by block, eigenvalues ascending, then the unit vectors).
 * S_out: p x p dense (group-block-diagonal); obj_out: the accumulated objective the reference returns. */
void kb_default_group_opts(kb_group_opts* opts);
int kb_solve_s_group(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, int32_t method, double m,
                     const kb_group_opts* opts, const double* V, int64_t nv, double* S_out, double* obj_out,
                     kb_group_info* info);
/* Multi-GPU form for a BLOCK-DIAGONAL Σ (BASELINE configs[3]): each rank passes its own diagonal
 * block(s) (no group straddles ranks).  All ranks sweep in lockstep and call `reduce` -- an all-reduce of a few
 * doubles (NCCL / MPI / anything) -- where the reference takes a GLOBAL decision: the shared equi γ
 * (min of the blocks' λmin, group.jl:490-491), the per-sweep convergence test on the summed objective and the
 * overall max δ (group.jl:1207-1212 ...), the SDP quick return (group.jl:982-984).  The union of the ranks'
 * S blocks equals the single-GPU solution of the whole matrix; obj_local_out is this shard's part of the objective. */
#define KB_REDUCE_MIN 0
#define KB_REDUCE_SUM 1
#define KB_REDUCE_MAX 2
typedef int (*kb_reduce_fn)(void* user, int32_t op, double* vals, int32_t n);   /* in place; 0 = ok */
int kb_solve_s_group_sharded(kb_ctx* ctx, const double* Sigma_block, int64_t p, const int64_t* groups, int32_t method,
                             double m, const kb_group_opts* opts, const double* V, int64_t nv, kb_reduce_fn reduce,
                             void* user, double* S_out, double* obj_local_out, kb_group_info* info);
/* `modelX_gaussian_group_knockoffs(X, method, groups, μ, Σ; m, kwargs...)` (group.jl:109-127):
 * S = solve_s_group(...), X̃ = condition(X, μ, Σ, S; m). */
int kb_modelx_gaussian_group_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* groups,
                                       const double* mu, const double* Sigma, int32_t method, int32_t m,
                                       const kb_group_opts* opts, const double* V, int64_t nv, const double* Z,
                                       uint64_t seed, int64_t row_offset, double* Xko_out, double* S_out,
                                       double* obj_out, kb_group_info* info);

/* ---- (2) Gram / covariance ------------------------------------------------------------------
 * G = (Xc' Xc) / denom with Xc = X - colmean if center != 0.  Replaces the Gram inside
 * `cov(covariance_approximator, X)` and `mean(X, dims=1)` (modelX.jl:47-49).
 * colmean_out may be NULL.  denom <= 0 means n.
 */
int kb_gram(kb_ctx* ctx, const double* X, int64_t n, int64_t p, int32_t center, double denom,
            double* G_out, double* colmean_out);
/* row-shard building block for multi-GPU: raw sums over the local rows, no centring:
 * dG (p x p, full symmetric) = X' X, dcolsum (p) = 1' X.  The caller all-reduces both (NCCL)
 * and calls kb_gram_finish_dev on every rank. */
int kb_gram_partial_dev(kb_ctx* ctx, const double* dX, int64_t n_local, int64_t ldx, int64_t p,
                        double* dG, double* dcolsum);
int kb_gram_finish_dev(kb_ctx* ctx, double* dG, double* dcolsum, int64_t n_total, int64_t p,
                       int32_t center, double denom);

/* Σ̂ = cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X) and μ̂ = mean(X, dims=1) -- the default
 * covariance_approximator of the 2-argument entry (modelX.jl:43-50; CovarianceEstimation.jl is not vendored:
 * restated from the published Ledoit-Wolf / Schäfer-Strimmer target-D formula).
 * shrinkage_out receives the intensity λ. */
int kb_cov_shrink_lw(kb_ctx* ctx, const double* X, int64_t n, int64_t p, double* Sigma_out, double* colmean_out,
                     double* shrinkage_out);

/* ---- (3) conditional-covariance factor ------------------------------------------------------
 * Replaces modelX.jl:106-120: ΣinvS = inv(Symmetric Σ)·S, C = 2S − S·ΣinvS and the Cholesky factor
 * used for the noise.  S is a length-p vector (Diagonal(s)) if s_is_diag, else dense p x p.
 * The pm x pm factor of modelX.jl:116-120 has exact block structure (all sub-diagonal blocks of
 * a block column are equal); it is returned as Lblocks = [L_11 | M_1 | L_22 | M_2 | ... | L_mm]
 * (2m-1 column-major p x p blocks): block row i, block col k of the dense factor is L_kk if
 * i == k, M_k if i > k.  For m == 1 Lblocks is just the lower factor L of C.
 */
int kb_cond_factor(kb_ctx* ctx, const double* Sigma, int64_t p, const double* S, int32_t s_is_diag,
                   int32_t m, double* SigmaInvS_out, double* Lblocks_out);

/* ---- (4) sampling -----------------------------------------------------------------------------
 * Xko (n x m·p) = repeat(X − (X .− μ')·ΣinvS, 1, m) + Z·L   (modelX.jl:112, 118-121; the noise
 * multiplies the LOWER factor exactly as the reference does).
 * Z (n x m·p, column-major) is the injected `randn(n, m·p)`; Z == NULL draws it on the device from
 * Philox4x32-10 keyed by (seed, row, column) so any row shard reproduces the same numbers.
 * row_offset = global index of X's first row (for row-sharded multi-GPU sampling).
 */
int kb_sample(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu,
              const double* SigmaInvS, const double* Lblocks, int32_t m, const double* Z,
              uint64_t seed, int64_t row_offset, double* Xko_out);

/* `condition(X, μ, Σ, S; m)` (modelX.jl:96-122) in one call: conditional factor + sampling with every intermediate
 * (Σ⁻¹S, the 2m−1 factor blocks) kept on the device.  S is a length-p vector (Diagonal(s), the single-knockoff path,
 * modelX.jl:64) if s_is_diag, else dense p x p (group knockoffs, group.jl:125).  X / Xko_out may be a ROW SHARD of the
 * data (n local rows, row_offset = global index of the first one): with the device RNG (Z == NULL) the union of the
 * shards equals the single call on all rows. */
int kb_condition(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu, const double* Sigma,
                 const double* S, int32_t s_is_diag, int32_t m, const double* Z, uint64_t seed, int64_t row_offset,
                 double* Xko_out);

/* ---- one-shot entry points ----------------------------------------------------------------------
 * kb_modelx_gaussian_knockoffs replaces `modelX_gaussian_knockoffs(X, method, μ, Σ; m, kwargs...)`
 * (modelX.jl:53-66): s = solve_s(...), Xko = condition(X, μ, Σ, Diagonal(s); m); all
 * intermediates stay on the device.  row_offset / n as in kb_sample.
 */
int kb_modelx_gaussian_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu,
                                 const double* Sigma, int32_t method, int32_t m,
                                 const kb_solve_opts* opts, const double* s_init, const double* Z,
                                 uint64_t seed, int64_t row_offset, double* Xko_out, double* s_out,
                                 kb_solve_info* info);

/* device-resident pieces of the same pipeline (used by the multi-GPU driver and bench.py) */
int kb_cond_factor_dev(kb_ctx* ctx, const double* dSigma, int64_t p, const double* dS, int32_t s_is_diag,
                       int32_t m, double* dSigmaInvS, double* dLblocks);
/* One block column k (1-based) of the factor for Diagonal(s), independent of the others: the recursion of
 * modelX.jl:116-120 has the closed form A_k = ((k+1)/k) S - (1/k) S (k Sigma - (k-1) S)^-1 S, L_kk = chol(A_k),
 * M_k = L_kk - S inv(L_kk)'.  Lets several devices / ranks build the m block columns side by side.
 * dSigmaInvS: written for k == 1 only (may be NULL); dMk: NULL allowed, not written for k == m. */
int kb_cond_factor_block_dev(kb_ctx* ctx, const double* dSigma, int64_t p, const double* ds, int32_t k, int32_t m,
                             double* dSigmaInvS, double* dLkk, double* dMk);
int kb_sample_dev(kb_ctx* ctx, const double* dX, int64_t n, int64_t ldx, int64_t p, const double* dmu,
                  const double* dSigmaInvS, const double* dLblocks, int32_t m, const double* dZ, int64_t ldz,
                  uint64_t seed, int64_t row_offset, double* dXko, int64_t ldxko);

/* ---- (5) feature statistics and the knockoff filter (the step right after sampling) ----
 * Importance scores of `fit_marginal` (fit_lasso.jl:214-222): with y_std = zscore(y), X_std / X̃_std the
 * column-wise z-scores (corrected std), T0 = abs2.(X_std' y_std) ./ n and Tk likewise for copy k.
 * T_out = [T0 | T1 | ... | Tm], (m+1)·p doubles.  One HBM-bound pass over X and X̃ (8·n·(m+1)·p bytes). */
int kb_marginal_stats(kb_ctx* ctx, const double* y, const double* X, const double* Xko, int64_t n, int64_t p,
                      int32_t m, double* T_out);
int kb_marginal_stats_dev(kb_ctx* ctx, const double* dy, const double* dX, int64_t ldx, const double* dXko,
                          int64_t ldxko, int64_t n, int64_t p, int32_t m, double* dT_out);
/* `MK_statistics(T0, Tk)` (threshold.jl:96-120 for m > 1: κ, τ, W; :134-142 for m == 1: W only, kappa_out /
 * tau_out untouched and may be NULL).  Tk = [T1 | ... | Tm] (m·p).  κ is the LAST k with |Tk| > |T0| and the
 * filter is the median, exactly as the reference computes them. */
int kb_mk_statistics(kb_ctx* ctx, const double* T0, const double* Tk, int64_t p, int32_t m, int64_t* kappa_out,
                     double* tau_out, double* W_out);
/* `threshold(w, q, method, rej_bounds)` (threshold.jl:19-31); knockoff_plus: 1 = :knockoff_plus, 0 = :knockoff;
 * reference default rej_bounds = 10000.  tau_out = +Inf when no threshold qualifies. */
int kb_threshold(kb_ctx* ctx, const double* w, int64_t p, double q, int32_t knockoff_plus, int64_t rej_bounds,
                 double* tau_out);
/* `mk_threshold(τ, κ, m, q, :knockoff_plus, rej_bounds)` (threshold.jl:55-75). */
int kb_mk_threshold(kb_ctx* ctx, const double* tau, const int64_t* kappa, int64_t p, int32_t m, double q,
                    int64_t rej_bounds, double* tau_hat_out);
/* `fit_marginal(y, ko; fdrs, filter_method)` (fit_lasso.jl:201-257) on an existing knockoff (X, Xko host
 * matrices): scores, optional group sums (groups != NULL: W has one entry per unique group, in order of first
 * appearance, fit_lasso.jl:224-238), MK_statistics, one threshold per target FDR and the selections.
 * Outputs (host; T_out, kappa_out, tau_out, selected_out, n_w_out may be NULL): T_out (m+1)·p, W_out n_w,
 * kappa_out / tau_out n_w (m > 1 only), taus_out nfdr, selected_out n_w x nfdr 0/1 mask (column f =
 * `findall(x -> x >= τ̂_f, W)`), n_w_out = length of W (p, or the number of groups). */
int kb_marginal_filter(kb_ctx* ctx, const double* y, const double* X, const double* Xko, int64_t n, int64_t p,
                       int32_t m, const int64_t* groups, const double* fdrs, int32_t nfdr, int32_t knockoff_plus,
                       double* T_out, double* W_out, int64_t* kappa_out, double* tau_out, double* taus_out,
                       int32_t* selected_out, int64_t* n_w_out);
/* `fit_marginal(y, X, μ, Σ; method, m, fdrs, filter_method)` (fit_lasso.jl:188-200 + 201-257), fused: the
 * knockoffs are generated chunk by chunk and their scores accumulated while each chunk is still on the
 * device, so X̃ is never stored nor copied back unless Xko_out != NULL (removes the n·m·p·8-byte D2H). */
int kb_fit_marginal(kb_ctx* ctx, const double* y, const double* X, int64_t n, int64_t p, const double* mu,
                    const double* Sigma, int32_t method, int32_t m, const kb_solve_opts* opts, const double* fdrs,
                    int32_t nfdr, int32_t knockoff_plus, const double* Z, uint64_t seed, double* Xko_out,
                    double* s_out, double* T_out, double* W_out, int64_t* kappa_out, double* tau_out,
                    double* taus_out, int32_t* selected_out, kb_solve_info* info);

/* ---- (6) approximate (block-diagonal covariance) knockoffs ---------------------------
 * `approx_modelX_gaussian_knockoffs(X, method, window_ranges; m, kwargs...)` (src/approx.jl:62-102).
 * window_offsets: nwin + 1 ascending 0-based column offsets, window w = columns [off[w], off[w+1]) (the
 * reference's `window_ranges`, which must tile 1:p in order for its BlockDiagonal / vcat assembly to make sense;
 * off[0] = 0 and off[nwin] = p are checked like approx.jl:73-75).  Per window: Σ̂_w = Ledoit-Wolf shrinkage
 * covariance of X[:, w] (approx.jl:84), s_w = solve_s(Σ̂_w, method; m) (:86); then γ = the `fzero` bisection of
 * :97 / :105-109 (the positive-definiteness probe is a Cholesky factorisation; γ = min over windows because Σ is
 * block diagonal), s .*= γ, and X̃ = condition(X, μ̂, BlockDiagonal(Σ̂_w), Diagonal(s); m) window by window.
 * Outputs: Xko_out n x m·p; s_out p; mu_out p (may be NULL); Sigma_blocks_out = the Σ̂_w packed one after the
 * other, window w column-major pw x pw (may be NULL); gamma_out (may be NULL). */
int kb_approx_modelx_gaussian_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p,
                                        const int64_t* window_offsets, int32_t nwin, int32_t method, int32_t m,
                                        const kb_solve_opts* opts, const double* Z, uint64_t seed, double* Xko_out,
                                        double* s_out, double* mu_out, double* Sigma_blocks_out, double* gamma_out,
                                        kb_solve_info* info);
/* The two phases of the call above for windows [w0, w1) only -- windows are independent, so ranks of a multi-GPU
 * job take disjoint window ranges, all-reduce(min) the γ they return and scale s before phase 2 (no other
 * exchange; X̃ columns of different windows are disjoint).  Arrays are full length (p, Σ pw²); only the entries
 * of the handled windows are read / written. */
int kb_approx_solve_windows(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* window_offsets,
                            int32_t nwin, int32_t w0, int32_t w1, int32_t method, int32_t m,
                            const kb_solve_opts* opts, double* Sigma_blocks_out, double* s_out, double* mu_out,
                            double* gamma_min_out, kb_solve_info* info);
int kb_approx_condition_windows(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* window_offsets,
                                int32_t nwin, int32_t w0, int32_t w1, int32_t m, const double* Sigma_blocks,
                                const double* s, const double* mu, const double* Z, uint64_t seed,
                                int64_t row_offset, double* Xko_out);
/* sup{γ in [0, 1] : (m+1)/m·Symmetric(Sigma) − γ·Diagonal(s) ≻ 0} -- approx.jl:97, 105-109 on one dense matrix. */
int kb_feasible_gamma(kb_ctx* ctx, const double* Sigma, int64_t p, const double* s, double m, double* gamma_out);

/* ---- (7) representative group knockoffs ---------------------------------------------
 * Index arrays at this boundary (group_reps, prioritize_idx) are 1-BASED like the reference's `Vector{Int}`.
 *
 * `choose_group_reps(Σ::Symmetric, groups; threshold, prioritize_idx)` (group.jl:1929-2033): Σ must be a correlation
 * matrix; prioritize_idx == NULL is `nothing`.  group_reps_out: capacity p, sorted; inv(Σ) runs on the device,
 * the per-group decisions are g x g host work (see rep.cu for the Schur-complement form of the ratio test). */
int kb_choose_group_reps(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, double threshold,
                         const int64_t* prioritize_idx, int64_t n_prioritize, int64_t* group_reps_out,
                         int64_t* n_reps_out);
/* `cond_indep_corr(Σ, groups, group_reps)` (group.jl:205-240): the covariance that satisfies the conditional
 * independence assumption exactly (p x p, symmetric). */
int kb_cond_indep_corr(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, const int64_t* group_reps,
                       int64_t n_reps, double* Sigma_new_out);
/* `solve_s_graphical_group(Σ::Symmetric, groups, group_reps, method; m, kwargs...)` (group.jl:268-303): the group
 * problem on the representatives (S_out, r x r), extended to all variables (D_out, p x p, |x| < 1e-10 zeroed,
 * symmetric).  V (r x nv, in the order of group_reps) as in kb_solve_s_group. */
int kb_solve_s_graphical_group(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups,
                               const int64_t* group_reps, int64_t n_reps, int32_t method, double m,
                               const kb_group_opts* opts, const double* V, int64_t nv, double* S_out, double* D_out,
                               double* obj_out, kb_group_info* info);
/* `modelX_gaussian_rep_group_knockoffs(X, method, groups, μ, Σ; m, rep_threshold, enforce_cond_indep, kwargs...)`
 * (group.jl:170-203).  group_reps_in == NULL: chosen by choose_group_reps(Σ, groups, threshold = rep_threshold).
 * Outputs: Xko_out n x m·p; group_reps_out (capacity p) / n_reps_out; S11_out r x r (capacity p·p is always enough);
 * D_out p x p (may be NULL); obj_out. */
int kb_modelx_gaussian_rep_group_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* groups,
                                           const double* mu, const double* Sigma, int32_t method, int32_t m,
                                           double rep_threshold, int32_t enforce_cond_indep, const kb_group_opts* opts,
                                           const int64_t* group_reps_in, int64_t n_reps_in, const double* V, int64_t nv,
                                           const double* Z, uint64_t seed, double* Xko_out, int64_t* group_reps_out,
                                           int64_t* n_reps_out, double* S11_out, double* D_out, double* obj_out,
                                           kb_group_info* info);

/* ---- (8) several GPUs behind one handle (SURVEY 8b `kb_create(ctx**, const int* devs, int ndev)`, 8e) -----------------
 * One host process, one kb_ctx and one host thread per listed device (a device may be listed more than once: two contexts
 * on it).  Only the naturally data-parallel pieces are partitioned; a single CCD solve is sequential and stays on one device.
 * All pointers are HOST pointers, as in the single-device calls above; results equal the single-device calls
 * (sampling: bit for bit with the device RNG or an injected Z; Gram: to summation order). */
typedef struct kb_multi kb_multi;
int kb_create_multi(kb_multi** mc, const int* devices, int32_t ndev);
void kb_destroy_multi(kb_multi* mc);
int32_t kb_multi_size(const kb_multi* mc);
kb_ctx* kb_multi_ctx(kb_multi* mc, int32_t i);              /* the i-th device's context (owned by mc) */
const char* kb_multi_last_error(const kb_multi* mc);
/* kb_gram with the rows of X sharded over the devices: per-device DMMA SYRK, then ONE all-reduce of the p x p sums done by
 * a peer-to-peer kernel over NVLink (each device reduces one slice of every peer's buffer and stores it back to all). */
int kb_multi_gram(kb_multi* mc, const double* X, int64_t n, int64_t p, int32_t center, double denom, double* G_out,
                  double* colmean_out);
/* kb_condition with the rows of X / Z / Xko sharded over the devices (no exchange: the RNG is keyed by the global row). */
int kb_multi_condition(kb_multi* mc, const double* X, int64_t n, int64_t p, const double* mu, const double* Sigma,
                       const double* S, int32_t s_is_diag, int32_t m, const double* Z, uint64_t seed, double* Xko_out);
/* kb_modelx_gaussian_knockoffs: the solve on device 0, `condition` row-sharded over all devices. */
int kb_multi_modelx_gaussian_knockoffs(kb_multi* mc, const double* X, int64_t n, int64_t p, const double* mu,
                                       const double* Sigma, int32_t method, int32_t m, const kb_solve_opts* opts,
                                       const double* s_init, const double* Z, uint64_t seed, double* Xko_out, double* s_out,
                                       kb_solve_info* info);
/* "replicas only": nprob independent solve_s problems on the same Sigma (methods[q], ms[q]) dealt round-robin to the devices
 * (configs[2]: :sdp_ccd and :mvr side by side).  s_out: p x nprob column-major; infos: nprob entries or NULL. */
int kb_multi_solve_s_many(kb_multi* mc, const double* Sigma, int64_t p, int32_t nprob, const int32_t* methods,
                          const double* ms, const kb_solve_opts* opts, double* s_out, kb_solve_info* infos);
/* solve_s_group on a BLOCK-DIAGONAL Sigma = blockdiag(Sigma_blocks[0..nblocks-1]) (no group straddles blocks; configs[3]):
 * block b on device b mod ndev, lockstep sweeps, the global decisions of kb_solve_s_group_sharded reduced in-process.
 * Sigma_blocks[b]: p_blocks[b] x p_blocks[b]; groups[b]: p_blocks[b] ids; S_out[b]: p_blocks[b] x p_blocks[b];
 * obj_out: the summed objective; infos: nblocks entries or NULL. */
int kb_multi_solve_s_group_blockdiag(kb_multi* mc, int32_t nblocks, const double* const* Sigma_blocks,
                                     const int64_t* p_blocks, const int64_t* const* groups, int32_t method, double m,
                                     const kb_group_opts* opts, double* const* S_out, double* obj_out, kb_group_info* infos);

/* ---- test hooks (exercised by tests/, not part of the reference interface) -------------------- */
/* C (M x N, col-major) = alpha * op(A) * op(B) + beta * C on the hand-written DMMA kernel.
 * transa / transb: 0 = 'N', 1 = 'T' (BLAS dgemm semantics, host pointers). */
int kb_test_dgemm(kb_ctx* ctx, int32_t transa, int32_t transb, int64_t M, int64_t N, int64_t K,
                  double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                  double* C, int64_t ldc);
/* the same on DEVICE pointers, enqueued on ctx's stream without synchronisation (kernel timing against cuBLAS) */
int kb_test_dgemm_dev(kb_ctx* ctx, int32_t transa, int32_t transb, int64_t M, int64_t N, int64_t K, double alpha,
                      const double* dA, int64_t lda, const double* dB, int64_t ldb, double beta, double* dC,
                      int64_t ldc);
/* measurement variant of the same product with TMA-staged operands (cp.async.bulk.tensor + mbarrier ring, 128-byte swizzle;
 * csrc/dgemm_tma.cu): C = alpha*A*B + beta*C, A M x K with m contiguous, B K x N with k contiguous, device pointers,
 * 16-byte aligned with even leading dimensions. */
int kb_test_dgemm_tma_dev(kb_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* dA, int64_t lda,
                          const double* dB, int64_t ldb, double beta, double* dC, int64_t ldc);
/* structure of factor blocks as the sampling stage detects it: band_out = 0 (every M_k = L_kk + upper triangular:
 * Diagonal(s)), a multiple of 16 up to 64 (banded below the diagonal: contiguous group blocks) or -1 (general form). */
int kb_test_factor_band(kb_ctx* ctx, const double* Lblocks, int64_t p, int32_t m, int32_t* band_out);
/* lower Cholesky factor (potrf) and inverse of an SPD matrix on the device kernels. */
int kb_test_potrf(kb_ctx* ctx, const double* A, int64_t p, double* L_out);
int kb_test_spd_inverse(kb_ctx* ctx, const double* A, int64_t p, double* Ainv_out);
/* the restated cholesky(Positive, Symmetric(A)) of csrc/positive.cu: L_out p x p lower, d_out p signs (+1 / -1 / 0). */
int kb_test_potrf_positive(kb_ctx* ctx, const double* A, int64_t p, double* L_out, double* d_out, int32_t* modified_out);
/* n x ncols block of the device Gaussian stream (what kb_sample uses when Z == NULL). */
int kb_test_randn(kb_ctx* ctx, uint64_t seed, int64_t row_offset, int64_t n, int64_t ncols, double* Z_out);

#ifdef __cplusplus
}
#endif
#endif /* KNOCKOFFS_B200_H */
