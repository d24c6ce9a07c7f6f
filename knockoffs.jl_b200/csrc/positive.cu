// cholesky(PositiveFactorizations.Positive, A) -- the factorisation `condition` uses (src/modelX.jl:111, :120) -- for the
// inputs on which it differs from the ordinary Cholesky factor: (numerically) singular or indefinite conditional
// covariances, e.g. the :equi solution (κΣ − S exactly singular) or an :sdp_ccd solution on the feasibility boundary.
//
// PARITY UNPINNED.  PositiveFactorizations.jl is a dependency that is NOT vendored under /root/reference; this file
// restates its published algorithm as recalled: an unpivoted LDLᵀ with D = diag(d), d_j ∈ {+1, −1, 0},
//     |pivot| > tol :  d_j = sign(pivot),  L_jj = sqrt(|pivot|),  L_ij = B_ij·d_j / L_jj,  B ← B − d_j ℓ ℓᵀ
//     otherwise     :  d_j = 0,            L_jj = 1,              L_ij = 0   (the column is dropped)
// with tol = 10·p·eps·max|diag(A)|, and the "positive" factor is L itself (i.e. D replaced by I).  For a positive definite A
// every d_j = +1 and L is the Cholesky factor, which is why the fast path (linalg.cu potrf_lower) is tried first and this
// one only runs when that fails AND the caller opted in (kb_set_positive_cholesky); otherwise the library keeps
// returning KB_ERR_NOT_PD.  Blocked right-looking, 64 columns per step; the trailing update is the DMMA GEMM.
#include "kb_common.cuh"
#include "linalg.cuh"
#include "positive.cuh"
#include <vector>

namespace kb {

constexpr int PNB = 64;

__global__ void maxabs_diag_kernel(const double* A, long ld, int p, double* out) {
    __shared__ double red[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) m = fmax(m, fabs(A[i + (long)i * ld]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// LDLᵀ with sign / zero handling of one nb x nb (<= 64) diagonal block, in place (lower; strict upper zeroed).
// d[0..nb): +1 / −1 / 0.  counts[0] += number of pivots with d != +1.
__global__ void __launch_bounds__(256, 1) positive_diag_kernel(double* A, long ld, int nb, const double* tolp, double tol_scale,
                                                               double* d, int* counts) {
    __shared__ double T[PNB][PNB + 1];
    __shared__ double sd[PNB];
    const int tid = threadIdx.x;
    const double tol = tolp[0] * tol_scale;
    for (int e = tid; e < PNB * PNB; e += blockDim.x) {
        const int i = e % PNB, k = e / PNB;
        T[i][k] = (i < nb && k < nb && i >= k) ? A[i + (long)k * ld] : 0.0;
    }
    __syncthreads();
    int modified = 0;
    for (int j = 0; j < nb; ++j) {
        const double Bjj = T[j][j];
        const bool keep = fabs(Bjj) > tol;             // uniform: every thread reads the same shared value
        const double dj = keep ? (Bjj > 0.0 ? 1.0 : -1.0) : 0.0;
        const double s = keep ? sqrt(fabs(Bjj)) : 1.0;
        if (dj != 1.0) ++modified;
        __syncthreads();
        if (tid == 0) { T[j][j] = s; sd[j] = dj; }
        // column j: ℓ_i = B_ij d_j / s   (0 for a dropped column)
        for (int i = j + 1 + tid; i < nb; i += blockDim.x) T[i][j] = keep ? T[i][j] * (dj / s) : 0.0;
        __syncthreads();
        if (keep) {
            // B_ik -= d_j ℓ_i ℓ_k on the lower triangle of the trailing block
            const int r = nb - j - 1;
            for (int e = tid; e < r * r; e += blockDim.x) {
                const int i = j + 1 + e % r, k = j + 1 + e / r;
                if (i >= k) T[i][k] -= dj * T[i][j] * T[k][j];
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += blockDim.x) {
        const int i = e % nb, k = e / nb;
        A[i + (long)k * ld] = (i >= k) ? T[i][k] : 0.0;
    }
    if (tid < nb) d[tid] = sd[tid];
    if (tid == 0 && modified) atomicAdd(counts, modified);
}

// Rows of X (rows x nb, ld ldx) solve  X D L11ᵀ = Bpanel  in place: x_j = (b_j − Σ_{k<j} x_k d_k L11[j,k]) / (d_j L11[j,j]),
// x_j = 0 where d_j = 0.  XD (optional, ld ldxd) receives x_j d_j.  One thread per row.
__global__ void positive_panel_kernel(double* X, long ldx, int rows, int nb, const double* __restrict__ L11, long ldl,
                                      const double* __restrict__ d, double* XD, long ldxd) {
    __shared__ double Ls[PNB][PNB + 1];
    __shared__ double sd[PNB];
    for (int e = threadIdx.x; e < PNB * PNB; e += blockDim.x) {
        const int i = e % PNB, k = e / PNB;
        Ls[i][k] = (i < nb && k < nb && i >= k) ? L11[i + (long)k * ldl] : 0.0;
    }
    if (threadIdx.x < PNB) sd[threadIdx.x] = threadIdx.x < nb ? d[threadIdx.x] : 0.0;
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double x[PNB];
#pragma unroll 8
    for (int j = 0; j < PNB; ++j) x[j] = (j < nb) ? X[r + (long)j * ldx] : 0.0;
    for (int j = 0; j < nb; ++j) {
        double acc = x[j];
        for (int k = 0; k < j; ++k) acc = fma(-x[k] * sd[k], Ls[j][k], acc);
        const double dj = sd[j];
        x[j] = (dj != 0.0) ? acc / (dj * Ls[j][j]) : 0.0;
    }
    for (int j = 0; j < nb; ++j) {
        X[r + (long)j * ldx] = x[j];
        if (XD) XD[r + (long)j * ldxd] = x[j] * sd[j];
    }
}

int potrf_positive_lower(kb_ctx* ctx, double* A, long ld, int p, double* d, int* modified_out) {
    Scope sc(ctx);
    double *tol, *XD;
    int* counts;
    const long ldx = (p + 1) & ~1L;
    KB_TRY(sc.alloc(&tol, 4));
    KB_TRY(sc.alloc(&counts, 4));
    KB_TRY(sc.alloc(&XD, (size_t)ldx * PNB));
    KB_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int), ctx->stream));
    maxabs_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, ld, p, tol);
    ctx->launches++;
    const double tol_scale = 10.0 * (double)p * 2.220446049250313e-16;    // default_δ(A) = 10·size(A,1)·eps
    for (int j0 = 0; j0 < p; j0 += PNB) {
        const int nb = (p - j0 < PNB) ? p - j0 : PNB;
        double* A11 = A + j0 + (long)j0 * ld;
        positive_diag_kernel<<<1, 256, 0, ctx->stream>>>(A11, ld, nb, tol, tol_scale, d + j0, counts);
        ctx->launches++;
        const int rem = p - j0 - nb;
        if (rem <= 0) break;
        double* A21 = A + (j0 + nb) + (long)j0 * ld;
        positive_panel_kernel<<<ceil_div(rem, 128), 128, 0, ctx->stream>>>(A21, ld, rem, nb, A11, ld, d + j0, XD, ldx);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
        // A22 −= (L21 D) L21ᵀ on the lower tiles
        double* A22 = A + (j0 + nb) + (long)(j0 + nb) * ld;
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, rem, nb, -1.0, XD, ldx, A21, ld, 1.0, A22, ld, 0, 1, 0));
    }
    KB_TRY(zero_strict_upper(ctx, A, ld, p));
    int h = 0;
    KB_CUDA(ctx, cudaMemcpyAsync(&h, counts, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (modified_out) *modified_out = h;
    return KB_OK;
}

// X (rows x p, ld ldx) ← X · L⁻ᵀ · D⁻¹ for the LDLᵀ factor (L lower with positive diagonal, d_j = 0 columns give 0):
// blocked forward substitution, 64 columns per step (GEMM update with the finished columns, then the panel kernel).
int trsm_right_ldlt(kb_ctx* ctx, double* X, long ldx, int rows, const double* L, long ldl, const double* d, int p) {
    Scope sc(ctx);
    double* XD;
    const long ldxd = (rows + 1) & ~1L;
    KB_TRY(sc.alloc(&XD, (size_t)ldxd * p));        // X·D of the finished columns
    for (int j0 = 0; j0 < p; j0 += PNB) {
        const int nb = (p - j0 < PNB) ? p - j0 : PNB;
        if (j0 > 0) {
            // X[:, J] −= (X D)[:, <J] · L[J, <J]ᵀ
            KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rows, nb, j0, -1.0, XD, ldxd, L + j0, ldl, 1.0, X + (long)j0 * ldx, ldx));
        }
        positive_panel_kernel<<<ceil_div(rows, 128), 128, 0, ctx->stream>>>(X + (long)j0 * ldx, ldx, rows, nb, L + j0 + (long)j0 * ldl, ldl,
                                                                            d + j0, XD + (long)j0 * ldxd, ldxd);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
    }
    return KB_OK;
}

// Y (ld ldy) = X · diag(d), rows x cols
__global__ void scale_cols_kernel(const double* X, long ldx, int rows, int cols, const double* d, double* Y, long ldy) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), k = (int)(idx / rows);
        Y[i + (long)k * ldy] = X[i + (long)k * ldx] * d[k];
    }
}

int scale_columns(kb_ctx* ctx, const double* X, long ldx, int rows, int cols, const double* d, double* Y, long ldy) {
    long g = ((long)rows * cols + 255) / 256, cap = (long)ctx->num_sms * 16;
    scale_cols_kernel<<<(int)(g < 1 ? 1 : (g > cap ? cap : g)), 256, 0, ctx->stream>>>(X, ldx, rows, cols, d, Y, ldy);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace kb
