// Feature-importance statistics and the knockoff filter on the device: the step right after
// sampling, so X̃ never has to leave HBM.  Reference: src/fit_lasso.jl:188-257 (fit_marginal),
// src/threshold.jl:19-31 (threshold), :55-75 (mk_threshold), :96-142 (MK_statistics).
#pragma once
#include "kb_common.cuh"

namespace kb {

// Per-column running sums of the one-pass standardised dot product, 4 doubles per column:
//   acc[4c + 0] = shift c0 (the column's first entry), [1] = Σ(x − c0), [2] = Σ(x − c0)², [3] = Σ(x − c0)·ys
constexpr int STATS_ACC = 4;

// y_std = zscore(y, mean(y), std(y)) (fit_lasso.jl:214); dys (n) receives it, d_sum_ys (1) its sum.
int zscore_vec_dev(kb_ctx* ctx, const double* dy, long n, double* dys, double* d_sum_ys);

// accumulate rows [0, rows) of the columns of A (rows x ncols, ld lda) against ys[0..rows):
// first != 0 initialises the accumulators (and takes the shift from row 0), else adds to them.
int marginal_accumulate_dev(kb_ctx* ctx, const double* dA, long rows, long lda, long ncols, const double* dys, double* dacc,
                            int first);
// T_c = abs2(X_std[:, c]' y_std) / n from the accumulators (fit_lasso.jl:217-222)
int marginal_finish_dev(kb_ctx* ctx, const double* dacc, long ncols, long n, const double* d_sum_ys, double* dT);

// group sums (fit_lasso.jl:224-238): Tg[k*ng + g] = Σ_{j: gidx[j] == g} |T[k*p + j]|, j ascending
int group_abs_sum_dev(kb_ctx* ctx, const double* dT, long p, int nblocks, const int* dgidx, int ng, double* dTg);

// MK_statistics (threshold.jl:96-120 for m > 1 -- κ = LAST k with |Tk| > |T0|, median of the rest --
// and :134-142 for m == 1: W = |T0| − |T1|; κ / τ are then not produced by the reference and are left untouched)
int mk_statistics_dev(kb_ctx* ctx, const double* dT0, const double* dTk, long p, int m, long long* dkappa, double* dtau,
                      double* dW);

// threshold(w, q, method, rej_bounds) (threshold.jl:19-31); offset = 0 (:knockoff) or 1 (:knockoff_plus).
// d_out[0] receives τ (+Inf if no t qualifies).
int threshold_dev(kb_ctx* ctx, const double* dw, long p, double q, int offset, long rej_bounds, double* d_out);
// mk_threshold(τ, κ, m, q, :knockoff_plus, rej_bounds) (threshold.jl:55-75)
int mk_threshold_dev(kb_ctx* ctx, const double* dtau, const long long* dkappa, long p, int m, double q, long rej_bounds,
                     double* d_out);
// sel[j] = (W[j] >= tau_hat) (fit_lasso.jl:250)
int select_dev(kb_ctx* ctx, const double* dW, long p, const double* d_tau_hat, int* dsel);

}  // namespace kb
