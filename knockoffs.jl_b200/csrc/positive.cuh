// cholesky(PositiveFactorizations.Positive, ·) restated (positive.cu): opt-in fallback of the conditional-factor stage.
#pragma once
#include "kb_common.cuh"

namespace kb {

// In-place LDLᵀ-with-sign factorisation of the lower triangle of A (p x p, ld); strict upper zeroed.
// d (device, p doubles): +1 / −1 / 0 per pivot.  modified_out: number of pivots with d != +1.
int potrf_positive_lower(kb_ctx* ctx, double* A, long ld, int p, double* d, int* modified_out);
// X (rows x p) ← X · L⁻ᵀ · D⁻¹ (columns with d_j = 0 become 0)
int trsm_right_ldlt(kb_ctx* ctx, double* X, long ldx, int rows, const double* L, long ldl, const double* d, int p);
// Y = X · diag(d)
int scale_columns(kb_ctx* ctx, const double* X, long ldx, int rows, int cols, const double* d, double* Y, long ldy);

}  // namespace kb
