// approx_modelX_gaussian_knockoffs (src/approx.jl:37-109) on the device -- see approx.cu.
#pragma once
#include "kb_common.cuh"

namespace kb {

int feasible_gamma_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* ds, double m, double* gamma_out);

int approx_solve_windows(kb_ctx* ctx, const double* dX, long n, long ldx, long p, const int64_t* win_off, int nwin, int w0,
                         int w1, int method, double m, const kb_solve_opts* opts, double* dSigblk, double* ds, double* dmu,
                         double* gamma_min_out, kb_solve_info* info);

int approx_condition_windows(kb_ctx* ctx, const double* dX, long n, long ldx, long p, const int64_t* win_off, int nwin,
                             int w0, int w1, int m, const double* dSigblk, const double* ds, const double* dmu,
                             const double* Z, uint64_t seed, int64_t row_offset, double* Xko);

int scale_vec_dev(kb_ctx* ctx, double* dx, long n, double a);

}  // namespace kb
