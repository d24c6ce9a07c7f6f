// Dense FP64 building blocks on the device, all built on the DMMA GEMM (dgemm_dmma.cuh):
// blocked Cholesky (potrf), triangular inverse (trtri), SPD inverse, symmetric helpers.
// Conventions: column-major, symmetric matrices carry their LOWER triangle; triangular factors
// are lower with an explicitly zeroed strict upper triangle (the GEMM's k-range clipping relies
// on real zeros there).
#pragma once
#include "kb_common.cuh"
#include "dgemm_dmma.cuh"

namespace kb {

constexpr int POTRF_NB = 128;

// generic GEMM launcher (device pointers).  AL/BL: layouts from dgemm_dmma.cuh.
int gemm(kb_ctx* ctx, int AL, int BL, const GemmParams& P);

// convenience: single-segment C = alpha*A*B + beta*C
int gemm1(kb_ctx* ctx, int AL, int BL, int M, int N, int K, double alpha, const double* A, long lda,
          const double* B, long ldb, double beta, double* C, long ldc, int seg_flags = 0, int tile_mode = 0,
          int store_mode = 0);

// In-place lower Cholesky of the lower triangle of A (p x p, ld).  Strict upper is zeroed.
// d_fail (device int, must be zero on entry) is set to 1 + (index of the failing pivot) if A is not PD.
// invdiag: >= ceil(p/POTRF_NB) * POTRF_NB^2 doubles; receives inv(L_bb) of every diagonal block
// (block b at invdiag + b*POTRF_NB^2, ld = POTRF_NB) -- reused by trtri_lower.
int potrf_lower(kb_ctx* ctx, double* A, long ld, int p, double* invdiag, int* d_fail);
static inline size_t potrf_invdiag_doubles(int p) {
    return (size_t)((p + POTRF_NB - 1) / POTRF_NB) * POTRF_NB * POTRF_NB;
}

// W = L^{-1} (lower, strict upper zeroed) for lower-triangular L.  W must not alias L.
// invdiag: the diagonal-block inverses produced by potrf_lower.  work: >= (p/2 + 2)^2 doubles.
int trtri_lower(kb_ctx* ctx, const double* L, long ldl, int p, const double* invdiag, double* W, long ldw,
                double* work);

// Ainv (lower triangle valid; full if symmetrize) = W' W
int syrk_WtW_lower(kb_ctx* ctx, const double* W, long ldw, int p, double* Out, long ldo);

// elementwise helpers
int zero_strict_upper(kb_ctx* ctx, double* A, long ld, int p);
int mirror_lower_to_upper(kb_ctx* ctx, double* A, long ld, int p);
int copy_matrix(kb_ctx* ctx, const double* A, long lda, double* B, long ldb, int rows, int cols);
int fill_zero(kb_ctx* ctx, double* A, size_t count);

}  // namespace kb
