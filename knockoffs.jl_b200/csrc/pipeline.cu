// Stages (2)-(4) on the device: Gram, conditional-covariance factor, fused sampling GEMM, RNG.
// Reference: src/modelX.jl:96-122 (condition), :39-51 (2-arg entry: cov + column means).
#include "pipeline.cuh"
#include "positive.cuh"
#include "linalg.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace kb {

static inline int ew_grid2(kb_ctx* ctx, long total, int threads = 256) {
    long g = (total + threads - 1) / threads;
    long cap = (long)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------
// counter-based Gaussian stream: Philox4x32-10 keyed by the seed, counter = (row, column pair).
// One Philox block -> two 53-bit uniforms -> Box-Muller pair -> columns (2c, 2c+1) of one row, so
// any row shard / any tile regenerates exactly the same Z (no state, no exchange between GPUs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void gaussian_pair(uint64_t seed, uint64_t row, uint32_t colpair, double& z0, double& z1) {
    uint32_t w[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), colpair, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const uint64_t a = ((uint64_t)w[0] << 32) | w[1];
    const uint64_t b = ((uint64_t)w[2] << 32) | w[3];
    const double u1 = (double)((a >> 11) + 1ULL) * 1.1102230246251565e-16;   // (0, 1]
    const double u2 = (double)(b >> 11) * 1.1102230246251565e-16;            // [0, 1)
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// Z (n x ncols, ld ldz); global row of local row i is row_offset + i; global column of local column c is col0 + c
__global__ void randn_kernel(uint64_t seed, int64_t row_offset, long n, long ncols, long col0, double* Z, long ldz) {
    const long npairs = (col0 + ncols + 1) / 2 - col0 / 2;
    const long total = n * npairs;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % n;
        const long cp = col0 / 2 + idx / n;
        double z0, z1;
        gaussian_pair(seed, (uint64_t)(row_offset + i), (uint32_t)cp, z0, z1);
        const long c0 = 2 * cp - col0, c1 = c0 + 1;
        if (c0 >= 0 && c0 < ncols) Z[i + c0 * ldz] = z0;
        if (c1 >= 0 && c1 < ncols) Z[i + c1 * ldz] = z1;
    }
}

int randn_dev(kb_ctx* ctx, uint64_t seed, int64_t row_offset, long n, long ncols, long col0, double* dZ, long ldz) {
    if (n <= 0 || ncols <= 0) return KB_OK;
    const long npairs = (col0 + ncols + 1) / 2 - col0 / 2;
    randn_kernel<<<ew_grid2(ctx, n * npairs), 256, 0, ctx->stream>>>(seed, row_offset, n, ncols, col0, dZ, ldz);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------
// conditional factor
// ------------------------------------------------------------------------------------------------
// Lower triangle (and zero strict upper) of Symmetric(Sigma) -- read from the UPPER triangle.
__global__ void sym_upper_to_lower_kernel(const double* Sig, long lds, int p, double* Out, long ldo) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bk = blockIdx.y;   // output tile rows bi, cols bk
    const int kr = bk * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bi * 32 + r;
        t[r][threadIdx.x] = (bi >= bk && kr < p && i < p && kr <= i) ? Sig[kr + (long)i * lds] : 0.0;
    }
    __syncthreads();
    const int i = bi * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = bk * 32 + r;
        if (i < p && k < p) Out[i + (long)k * ldo] = (i >= k) ? t[threadIdx.x][r] : 0.0;
    }
}

// Full symmetric matrix from the UPPER triangle (both triangles written).
__global__ void sym_upper_to_full_kernel(const double* Sig, long lds, int p, double* Out, long ldo) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        Out[i + (long)k * ldo] = (i <= k) ? Sig[i + (long)k * lds] : Sig[k + (long)i * lds];
    }
}

// SiS = Sinv * Diagonal(s);  C = 2 Diagonal(s) - Diagonal(s) * SiS      (modelX.jl:107-108)
__global__ void sis_c_diag_kernel(const double* Sinv, long ldi, const double* s, int p, double* SiS, long ldo,
                                  double* C, long ldc) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        const double v = Sinv[i + (long)k * ldi] * s[k];
        SiS[i + (long)k * ldo] = v;
        C[i + (long)k * ldc] = ((i == k) ? 2.0 * s[i] : 0.0) - s[i] * v;
    }
}
// B = C - S (S diagonal vector or dense p x p with ld lds)
__global__ void sub_S_kernel(const double* C, long ldc, const double* S, long lds, int s_is_diag, int p, double sign,
                             double* B, long ldb) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        const double sv = s_is_diag ? ((i == k) ? S[i] : 0.0) : S[i + (long)k * lds];
        B[i + (long)k * ldb] = C[i + (long)k * ldc] + sign * sv;
    }
}
__global__ void negate_kernel(const double* A, long lda, int rows, int cols, double* B, long ldb) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), k = (int)(idx / rows);
        B[i + (long)k * ldb] = -A[i + (long)k * lda];
    }
}
// symmetrize from the upper triangle in place: A(i,k) = A(k,i) for i > k  (Symmetric(C), uplo = :U)
__global__ void symmetrize_from_upper_kernel(double* A, long ld, int p) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bk = blockIdx.y;   // writes tile (rows bi, cols bk), bi >= bk
    if (bi < bk) return;
    const int kr = bk * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bi * 32 + r;
        t[r][threadIdx.x] = (kr < p && i < p) ? A[kr + (long)i * ld] : 0.0;   // element (kr, i) of the upper part
    }
    __syncthreads();
    const int i = bi * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = bk * 32 + r;
        if (i < p && k < p && i > k) A[i + (long)k * ld] = t[threadIdx.x][r];
    }
}

// Diagonal S: M_k = B_k L_kk^{-T} with B_k = A_k − S = L_kk L_kk' − S  =>  M_k = L_kk − S · W' where W = L_kk^{-1}
// (elementwise, W read transposed through a shared-memory tile).
__global__ void mk_from_factor_kernel(const double* __restrict__ L, long ldl, const double* __restrict__ W, long ldw,
                                      const double* __restrict__ s, int p, double* __restrict__ M, long ldm) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;          // output tile rows bi, cols bj  <-  W tile rows bj, cols bi
    const int wr = bj * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int wc = bi * 32 + r;
        t[r][threadIdx.x] = (wr < p && wc < p) ? W[wr + (long)wc * ldw] : 0.0;      // W(j, i)
    }
    __syncthreads();
    const int i = bi * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = bj * 32 + r;
        if (i < p && j < p) M[i + (long)j * ldm] = L[i + (long)j * ldl] - s[i] * t[threadIdx.x][r];
    }
}
// A_{k+1} = B_{k+1} + S with B_{k+1} = B_k − M_k M_k' = S − S A_k^{-1} S   =>   A_{k+1} = 2S − S A_k^{-1} S
__global__ void next_A_diag_kernel(const double* __restrict__ Ainv, long ldi, const double* __restrict__ s, int p,
                                   double* __restrict__ C, long ldc) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        C[i + (long)k * ldc] = ((i == k) ? 2.0 * s[i] : 0.0) - s[i] * Ainv[i + (long)k * ldi] * s[k];
    }
}

// C = 2S − S·(Σ⁻¹S) for diagonal S from the stored Σ⁻¹S (used when the conditional-factor stage has to restart)
__global__ void next_A_from_SiS_kernel(const double* __restrict__ SiS, long lds, const double* __restrict__ s, int p,
                                       double* __restrict__ C, long ldc) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        C[i + (long)k * ldc] = ((i == k) ? 2.0 * s[i] : 0.0) - s[i] * SiS[i + (long)k * lds];
    }
}

// T_k = k Σ − (k − 1) S for diagonal S, in place on the full symmetric copy of Σ
__global__ void block_T_kernel(double* __restrict__ A, long ld, const double* __restrict__ s, int p, double k) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), j = (int)(idx / p);
        const double v = k * A[i + (long)j * ld];
        A[i + (long)j * ld] = (i == j) ? v - (k - 1.0) * s[i] : v;
    }
}
// A_k = ((k + 1) / k) S − (1 / k) S T_k⁻¹ S for diagonal S
__global__ void block_A_kernel(const double* __restrict__ Tinv, long ldi, const double* __restrict__ s, int p, double k,
                               double* __restrict__ C, long ldc) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    const double rk = 1.0 / k;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), j = (int)(idx / p);
        C[i + (long)j * ldc] = ((i == j) ? (k + 1.0) * rk * s[i] : 0.0) - rk * s[i] * Tinv[i + (long)j * ldi] * s[j];
    }
}

static int check_fail(kb_ctx* ctx, int* d_fail, const char* what) {
    int f = 0;
    KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f != 0)
        return set_error(ctx, KB_ERR_NOT_PD, "PosDefException: %s is not positive definite; Cholesky factorization "
                                             "failed at pivot %d.", what, f);
    return KB_OK;
}

// SPD inverse (full symmetric output) of the matrix whose lower triangle is in A (destroyed).
static int spd_inverse_full(kb_ctx* ctx, double* A, long ld, int p, double* invdiag, int* d_fail, double* W,
                            double* work, double* Out, long ldo, const char* what) {
    KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
    KB_TRY(potrf_lower(ctx, A, ld, p, invdiag, d_fail));
    KB_TRY(check_fail(ctx, d_fail, what));
    KB_TRY(trtri_lower(ctx, A, ld, p, invdiag, W, ld, work));
    KB_TRY(syrk_WtW_lower(ctx, W, ld, p, Out, ldo));
    return mirror_lower_to_upper(ctx, Out, ldo, p);
}

int spd_inverse_dev(kb_ctx* ctx, const double* dA_upper, long lda, int p, double* dOut, long ldo) {
    const long ld = (p + 1) & ~1L;
    Scope sc(ctx);
    double *A, *W, *work, *invdiag;
    int* d_fail;
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&W, (size_t)ld * p));
    KB_TRY(sc.alloc(&work, (size_t)ld * p / 2 + 2 * ld + 64));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&d_fail, 4));
    dim3 grid(ceil_div(p, 32), ceil_div(p, 32)), block(32, 8);
    sym_upper_to_lower_kernel<<<grid, block, 0, ctx->stream>>>(dA_upper, lda, p, A, ld);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return spd_inverse_full(ctx, A, ld, p, invdiag, d_fail, W, work, dOut, ldo, "matrix");
}

int potrf_dev(kb_ctx* ctx, const double* dA_upper, long lda, int p, double* dL, long ldl) {
    Scope sc(ctx);
    double* invdiag;
    int* d_fail;
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&d_fail, 4));
    dim3 grid(ceil_div(p, 32), ceil_div(p, 32)), block(32, 8);
    sym_upper_to_lower_kernel<<<grid, block, 0, ctx->stream>>>(dA_upper, lda, p, dL, ldl);
    ctx->launches++;
    KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
    KB_TRY(potrf_lower(ctx, dL, ldl, p, invdiag, d_fail));
    return check_fail(ctx, d_fail, "matrix");
}

// modelX.jl:106-120.  Output blocks (each p x p, ld ldo, block stride ldo*p):
//   dLb = [L_11 | M_1 | L_22 | M_2 | ... | L_mm]   (2m-1 blocks),  L_kk lower with zero strict upper.
// Structure of the pm x pm factor: with A_1 = C, B_1 = C − S,
//   L_kk = chol(A_k),  M_k = B_k L_kk^{-T},  B_{k+1} = B_k − M_k M_k',  A_{k+1} = B_{k+1} + S.
// ONE block column of the pm x pm factor of modelX.jl:116-120 for Diagonal(s), WITHOUT the recursion over the earlier ones.
// The diagonal blocks of the factor are the Cholesky factors of A_1 = C, A_{k+1} = 2S − S A_k⁻¹ S (cond_factor_dev), and that
// recursion has the closed form
//     A_k = ((k + 1) / k) S − (1 / k) S (k Σ − (k − 1) S)⁻¹ S          (k = 1: 2S − S Σ⁻¹ S)
// (induction on k with the Woodbury identity), so the m block columns are independent: L_kk = chol(A_k),
// M_k = L_kk − S inv(L_kk)ᵀ.  k Σ − (k − 1) S = (k − 1)((k / (k − 1)) Σ − S) is at least Σ / m away from the boundary of the
// feasible set for every k <= m, so this form does not invert the nearly singular A_k an :sdp_ccd solution produces.
// One rank / device per k: 5/3 p³ flops each (4/3 for k = m) instead of 5.3 p³ in sequence at m = 5.
// dSiS (k == 1 only, may be null) receives Σ⁻¹S; dMk may be null (and is not formed for k == m).
int cond_factor_block_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* ds, int k, int m,
                          double* dSiS, double* dLkk, double* dMk, long ldo) {
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    if (k < 1 || k > m || p < 1) return set_error(ctx, KB_ERR_INVALID, "cond_factor_block: k must be in 1..m");
    const long ld = (p + 1) & ~1L;
    Scope sc(ctx);
    double *A, *W, *work, *invdiag, *Tinv, *C;
    int* d_fail;
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&W, (size_t)ld * p));
    KB_TRY(sc.alloc(&work, (size_t)ld * p / 2 + 2 * ld + 64));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&Tinv, (size_t)ld * p));
    KB_TRY(sc.alloc(&C, (size_t)ld * p));
    KB_TRY(sc.alloc(&d_fail, 4));
    dim3 tgrid(ceil_div(p, 32), ceil_div(p, 32)), tblock(32, 8);
    const int eg = ew_grid2(ctx, (long)p * p);
    sym_upper_to_lower_kernel<<<tgrid, tblock, 0, ctx->stream>>>(dSigma, lds, p, A, ld);
    ctx->launches++;
    if (k > 1) {       // only the lower triangle is read by the factorisation
        block_T_kernel<<<eg, 256, 0, ctx->stream>>>(A, ld, ds, p, (double)k);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(spd_inverse_full(ctx, A, ld, p, invdiag, d_fail, W, work, Tinv, ld, k == 1 ? "Sigma" : "k*Sigma - (k-1)*S"));
    if (k == 1) {
        // Σ⁻¹S and A_1 = C = 2S − S Σ⁻¹S in one pass; without an output for Σ⁻¹S it lands in the scratch matrix A
        sis_c_diag_kernel<<<eg, 256, 0, ctx->stream>>>(Tinv, ld, ds, p, dSiS ? dSiS : A, dSiS ? ldo : ld, C, ld);
    } else {
        block_A_kernel<<<eg, 256, 0, ctx->stream>>>(Tinv, ld, ds, p, (double)k, C, ld);
    }
    ctx->launches++;
    symmetrize_from_upper_kernel<<<tgrid, tblock, 0, ctx->stream>>>(C, ld, p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(copy_matrix(ctx, C, ld, dLkk, ldo, p, p));
    KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
    KB_TRY(potrf_lower(ctx, dLkk, ldo, p, invdiag, d_fail));
    KB_TRY(check_fail(ctx, d_fail, "the conditional covariance of the multiple-knockoff model"));
    if (k < m && dMk) {
        KB_TRY(trtri_lower(ctx, dLkk, ldo, p, invdiag, W, ld, work));
        mk_from_factor_kernel<<<tgrid, tblock, 0, ctx->stream>>>(dLkk, ldo, W, ld, ds, p, dMk, ldo);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
    }
    return KB_OK;
}

int cond_factor_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* dS, long ldS, int s_is_diag,
                    int m, double* dSiS, double* dLb, long ldo) {
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);   // modelX.jl:105
    if (p < 1) return set_error(ctx, KB_ERR_INVALID, "p must be >= 1");
    const long ld = (p + 1) & ~1L;
    Scope sc(ctx);
    double *A, *W, *work, *invdiag, *Sinv, *C, *B, *Sfull = nullptr;
    int* d_fail;
    KB_TRY(sc.alloc(&A, (size_t)ld * p));
    KB_TRY(sc.alloc(&W, (size_t)ld * p));
    KB_TRY(sc.alloc(&work, (size_t)ld * p / 2 + 2 * ld + 64));
    KB_TRY(sc.alloc(&invdiag, potrf_invdiag_doubles(p)));
    KB_TRY(sc.alloc(&Sinv, (size_t)ld * p));
    KB_TRY(sc.alloc(&C, (size_t)ld * p));
    KB_TRY(sc.alloc(&d_fail, 4));
    dim3 tgrid(ceil_div(p, 32), ceil_div(p, 32)), tblock(32, 8);
    const int eg = ew_grid2(ctx, (long)p * p);

    // Σinv = inv(Symmetric(Σ))                                             (modelX.jl:106)
    sym_upper_to_lower_kernel<<<tgrid, tblock, 0, ctx->stream>>>(dSigma, lds, p, A, ld);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    KB_TRY(spd_inverse_full(ctx, A, ld, p, invdiag, d_fail, W, work, Sinv, ld, "Sigma"));

    // ΣinvS = Σinv * S ;  C = 2S − S * ΣinvS                               (modelX.jl:107-108)
    if (s_is_diag) {
        sis_c_diag_kernel<<<eg, 256, 0, ctx->stream>>>(Sinv, ld, dS, p, dSiS, ldo, C, ld);
        ctx->launches++;
    } else {
        // S arrives dense; the reference multiplies the full matrices (group.jl:125 -> modelX.jl:107-108)
        KB_TRY(sc.alloc(&Sfull, (size_t)ld * p));
        KB_TRY(copy_matrix(ctx, dS, ldS, Sfull, ld, p, p));
        KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, p, p, p, 1.0, Sinv, ld, Sfull, ld, 0.0, dSiS, ldo));
        // C = 2S − S * SiS
        GemmParams P{};
        P.M = p; P.N = p; P.nseg = 1;
        P.seg[0] = GemmSeg{Sfull, ld, dSiS, ldo, p, 0};
        P.C = C; P.ldc = ld; P.Cin = Sfull; P.ldcin = ld; P.alpha = -1.0; P.beta = 2.0;
        KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, P));
    }
    KB_CUDA(ctx, cudaGetLastError());
    // Symmetric(C): the upper triangle is the one the reference trusts (modelX.jl:110, :120)
    symmetrize_from_upper_kernel<<<tgrid, tblock, 0, ctx->stream>>>(C, ld, p);
    ctx->launches++;

    const size_t bstride = (size_t)ldo * p;
    ctx->positive_modified = 0;
    if (m == 1) {
        KB_TRY(copy_matrix(ctx, C, ld, dLb, ldo, p, p));
        KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
        KB_TRY(potrf_lower(ctx, dLb, ldo, p, invdiag, d_fail));
        if (!ctx->positive_cholesky) return check_fail(ctx, d_fail, "the conditional covariance 2S - S*inv(Sigma)*S");
        int f = 0;
        KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (f == 0) return KB_OK;
        // cholesky(Positive, Symmetric(C)) (modelX.jl:111), restated: see positive.cu
        double* dsign;
        KB_TRY(sc.alloc(&dsign, (size_t)p));
        KB_TRY(copy_matrix(ctx, C, ld, dLb, ldo, p, p));
        return potrf_positive_lower(ctx, dLb, ldo, p, dsign, &ctx->positive_modified);
    }
    // Every factorisation below is an ordinary Cholesky; when one of them fails and the caller opted in
    // (kb_set_positive_cholesky), the whole block recursion is redone with the restated cholesky(Positive, ·) of the
    // pm x pm matrix (modelX.jl:120) in its LDLᵀ block form -- cond_factor_positive() at the end of this function.
    auto fail_or_positive = [&](const char* what) -> int {
        int f = 0;
        KB_CUDA(ctx, cudaMemcpyAsync(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (f == 0) return KB_OK;
        if (!ctx->positive_cholesky)
            return set_error(ctx, KB_ERR_NOT_PD, "PosDefException: %s is not positive definite; Cholesky factorization "
                                                 "failed at pivot %d.", what, f);
        return 1;     // > 0: redo with the Positive restatement
    };
    bool redo_positive = false;
    if (s_is_diag) {
        // Diagonal S (the single-knockoff path): with W = L_kk^{-1},
        //     M_k = (L_kk L_kk' − S) L_kk^{-T} = L_kk − S W'                 (elementwise)
        //     A_{k+1} = S − S A_k^{-1} S + S = 2S − S (W'W) S                 (one SYRK + elementwise)
        // -- the same blocks as the recursion B_{k+1} = B_k − M_k M_k' below, without its two p^3 GEMMs per copy
        // (5.3 p^3 flops executed instead of 12 p^3 at m = 5).  Sinv is free again and holds A_k^{-1}.
        for (int k = 0; k < m; ++k) {
            double* Lkk = dLb + (size_t)(2 * k) * bstride;
            KB_TRY(copy_matrix(ctx, C, ld, Lkk, ldo, p, p));
            KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
            KB_TRY(potrf_lower(ctx, Lkk, ldo, p, invdiag, d_fail));
            {
                const int rc = fail_or_positive("the conditional covariance of the multiple-knockoff model");
                if (rc < 0) return rc;
                if (rc > 0) { redo_positive = true; break; }
            }
            if (k == m - 1) break;
            double* Mk = dLb + (size_t)(2 * k + 1) * bstride;
            KB_TRY(trtri_lower(ctx, Lkk, ldo, p, invdiag, W, ld, work));
            mk_from_factor_kernel<<<tgrid, tblock, 0, ctx->stream>>>(Lkk, ldo, W, ld, dS, p, Mk, ldo);
            ctx->launches++;
            KB_TRY(syrk_WtW_lower(ctx, W, ld, p, Sinv, ld));
            KB_TRY(mirror_lower_to_upper(ctx, Sinv, ld, p));
            next_A_diag_kernel<<<eg, 256, 0, ctx->stream>>>(Sinv, ld, dS, p, C, ld);
            ctx->launches++;
        }
        KB_CUDA(ctx, cudaGetLastError());
        if (!redo_positive) return KB_OK;
    }
    KB_TRY(sc.alloc(&B, (size_t)ld * p));
    const double* Sptr = s_is_diag ? dS : Sfull;
    const long Sld = s_is_diag ? 0 : ld;
    if (redo_positive) {
        // the diagonal-S recursion overwrote C with a later A_k: rebuild A_1 = C = 2S − S Σ⁻¹S (Σ⁻¹S is still in dSiS)
        next_A_from_SiS_kernel<<<eg, 256, 0, ctx->stream>>>(dSiS, ldo, dS, p, C, ld);
        ctx->launches++;
        symmetrize_from_upper_kernel<<<tgrid, tblock, 0, ctx->stream>>>(C, ld, p);
        ctx->launches++;
    }
    sub_S_kernel<<<eg, 256, 0, ctx->stream>>>(C, ld, Sptr, Sld, s_is_diag, p, -1.0, B, ld);   // B_1 = C − S
    ctx->launches++;
    // A_1 = C (already in C); C doubles as the A_k buffer
    for (int k = 0; k < m && !redo_positive; ++k) {
        double* Lkk = dLb + (size_t)(2 * k) * bstride;
        KB_TRY(copy_matrix(ctx, C, ld, Lkk, ldo, p, p));
        KB_CUDA(ctx, cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
        KB_TRY(potrf_lower(ctx, Lkk, ldo, p, invdiag, d_fail));
        {
            const int rc = fail_or_positive("the conditional covariance of the multiple-knockoff model");
            if (rc < 0) return rc;
            if (rc > 0) { redo_positive = true; break; }
        }
        if (k == m - 1) break;
        double* Mk = dLb + (size_t)(2 * k + 1) * bstride;
        KB_TRY(trtri_lower(ctx, Lkk, ldo, p, invdiag, W, ld, work));
        // M_k = B_k W' = (L_kk L_kk' − S) W' = L_kk − S W'   :  B(kk, n) = W(n, kk), zero for kk > n.  Same p^3 product as
        // B_k W', but where S W' is structurally zero (S group-block-diagonal, W lower: rows more than g − 1 below the
        // diagonal) the result is L_kk BIT FOR BIT, which is what lets the sampling stage use two (banded-)triangular
        // products per copy (sample_plan_create).
        {
            GemmParams P{};
            P.M = p; P.N = p; P.nseg = 1;
            P.seg[0] = GemmSeg{Sfull, ld, W, ld, p, SEG_KHI_N};
            P.C = Mk; P.ldc = ldo; P.Cin = Lkk; P.ldcin = ldo; P.alpha = -1.0; P.beta = 1.0;
            KB_TRY(gemm(ctx, A_MCONTIG, B_NCONTIG, P));
        }
        // B_{k+1} = B_k − M_k M_k'  (symmetric: lower tiles only -- half the flops -- then mirrored)
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, p, p, p, -1.0, Mk, ldo, Mk, ldo, 1.0, B, ld, 0, 1, 0));
        KB_TRY(mirror_lower_to_upper(ctx, B, ld, p));
        // A_{k+1} = B_{k+1} + S
        sub_S_kernel<<<eg, 256, 0, ctx->stream>>>(B, ld, Sptr, Sld, s_is_diag, p, +1.0, C, ld);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaGetLastError());
    if (!redo_positive) return KB_OK;

    // ---- cholesky(Positive, Symmetric(Σ̃)) of the pm x pm matrix (modelX.jl:116-120), restated (positive.cu) in block
    //      form: with A_1 = C, B_1 = C − S and, per block column k, the LDLᵀ-with-sign factor (L_kk, D_k) of A_k,
    //          M_k = B_k L_kk⁻ᵀ D_k⁻¹,   B_{k+1} = B_k − M_k D_k M_kᵀ,   A_{k+1} = B_{k+1} + S
    //      (every sub-diagonal block of block column k equals M_k, exactly as in the ordinary factor).
    if (!s_is_diag) {
        // dense S: C was consumed by the loop above; rebuild A_1 = 2S − S·(Σ⁻¹S)
        GemmParams P{};
        P.M = p; P.N = p; P.nseg = 1;
        P.seg[0] = GemmSeg{Sfull, ld, dSiS, ldo, p, 0};
        P.C = C; P.ldc = ld; P.Cin = Sfull; P.ldcin = ld; P.alpha = -1.0; P.beta = 2.0;
        KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, P));
        symmetrize_from_upper_kernel<<<tgrid, tblock, 0, ctx->stream>>>(C, ld, p);
        ctx->launches++;
    } else {
        next_A_from_SiS_kernel<<<eg, 256, 0, ctx->stream>>>(dSiS, ldo, dS, p, C, ld);
        ctx->launches++;
        symmetrize_from_upper_kernel<<<tgrid, tblock, 0, ctx->stream>>>(C, ld, p);
        ctx->launches++;
    }
    sub_S_kernel<<<eg, 256, 0, ctx->stream>>>(C, ld, Sptr, Sld, s_is_diag, p, -1.0, B, ld);
    ctx->launches++;
    double *dsign, *XD;
    KB_TRY(sc.alloc(&dsign, (size_t)p));
    KB_TRY(sc.alloc(&XD, (size_t)ld * p));
    int modified_total = 0;
    for (int k = 0; k < m; ++k) {
        double* Lkk = dLb + (size_t)(2 * k) * bstride;
        KB_TRY(copy_matrix(ctx, C, ld, Lkk, ldo, p, p));
        int modified = 0;
        KB_TRY(potrf_positive_lower(ctx, Lkk, ldo, p, dsign, &modified));
        modified_total += modified;
        if (k == m - 1) break;
        double* Mk = dLb + (size_t)(2 * k + 1) * bstride;
        KB_TRY(copy_matrix(ctx, B, ld, Mk, ldo, p, p));
        KB_TRY(trsm_right_ldlt(ctx, Mk, ldo, p, Lkk, ldo, dsign, p));
        KB_TRY(scale_columns(ctx, Mk, ldo, p, p, dsign, XD, ld));
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, p, p, p, -1.0, XD, ld, Mk, ldo, 1.0, B, ld, 0, 1, 0));
        KB_TRY(mirror_lower_to_upper(ctx, B, ld, p));
        sub_S_kernel<<<eg, 256, 0, ctx->stream>>>(B, ld, Sptr, Sld, s_is_diag, p, +1.0, C, ld);
        ctx->launches++;
    }
    ctx->positive_modified = modified_total;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------
// sampling
// ------------------------------------------------------------------------------------------------
// bias_n = sum_k mu_k SiS(k, n)      (the "− μ'" part of (X .− μ') ΣinvS, modelX.jl:112)
__global__ void mu_sis_kernel(const double* mu, const double* SiS, long ld, int p, double* bias) {
    const int n = blockIdx.x;
    __shared__ double red[256];
    double acc = 0.0;
    for (int k = threadIdx.x; k < p; k += blockDim.x) acc = fma(mu[k], SiS[k + (long)n * ld], acc);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) bias[n] = red[0];
}
// Suffix sums of the noise blocks: T_k = Σ_{i>k} Z_i for k = 0 .. m−2 (T_k at T + k * tstride, Z_i at Z + i * zstride),
// one pass (reads m−1 blocks, writes m−1 blocks).
__global__ void suffix_sum_kernel(double* __restrict__ T, long ldt, long tstride, const double* __restrict__ Z, long ldz,
                                  long zstride, long rows, int cols, int m) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, c = idx / rows;
        double acc = 0.0;
        for (int k = m - 2; k >= 0; --k) {
            acc += Z[i + c * ldz + (long)(k + 1) * zstride];
            T[i + c * ldt + (long)k * tstride] = acc;
        }
    }
}

// Inclusive suffix sums S_k = Σ_{i>=k} Z_i for k = 0 .. m−1 (S_k at T + k * tstride).
__global__ void suffix_sum_incl_kernel(double* __restrict__ T, long ldt, long tstride, const double* __restrict__ Z, long ldz,
                                       long zstride, long rows, int cols, int m) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, c = idx / rows;
        double acc = 0.0;
        for (int k = m - 1; k >= 0; --k) {
            acc += Z[i + c * ldz + (long)k * zstride];
            T[i + c * ldt + (long)k * tstride] = acc;
        }
    }
}
// band = max over k and over entries i > j with M_k(i,j) != L_kk(i,j) of (i − j)   (blocks [L_11 | M_1 | L_22 | ...]);
// 0 if the strict lower triangles agree bit for bit.
__global__ void factor_structure_kernel(const double* __restrict__ Lb, long ldo, long bstride, int p, int m, int* __restrict__ band) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)p * p, total = per * (m - 1);
    int bw = 0;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int k = (int)(idx / per);
        const long e = idx % per;
        const int i = (int)(e % p), j = (int)(e / p);
        if (i > j) {
            const double l = Lb[(size_t)(2 * k) * bstride + i + (long)j * ldo], mk = Lb[(size_t)(2 * k + 1) * bstride + i + (long)j * ldo];
            if (!(l == mk)) bw = max(bw, i - j);
        }
    }
    if (bw) atomicMax(band, bw);
}
// N_k = M_k − L_kk on and above the band (i <= j + band), zero below it
__global__ void factor_upper_part_kernel(const double* __restrict__ Lb, long ldo, long bstride, int p, int m, int band,
                                         double* __restrict__ N, long ldn) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)p * p, total = per * (m - 1);
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int k = (int)(idx / per);
        const long e = idx % per;
        const int i = (int)(e % p), j = (int)(e / p);
        double v = 0.0;
        if (i <= j + band) {
            v = Lb[(size_t)(2 * k + 1) * bstride + i + (long)j * ldo];
            if (i >= j) v -= Lb[(size_t)(2 * k) * bstride + i + (long)j * ldo];
        }
        N[(size_t)k * ldn * p + i + (long)j * ldn] = v;
    }
}

int factor_band_dev(kb_ctx* ctx, const double* dLb, long ldo, int p, int m, int* band_out) {
    *band_out = 0;
    if (m <= 1) return KB_OK;
    Scope sc(ctx);
    int* dflag;
    KB_TRY(sc.alloc(&dflag, 4));
    KB_CUDA(ctx, cudaMemsetAsync(dflag, 0, sizeof(int), ctx->stream));
    factor_structure_kernel<<<ew_grid2(ctx, (long)p * p * (m - 1)), 256, 0, ctx->stream>>>(dLb, ldo, (long)ldo * p, p, m, dflag);
    ctx->launches++;
    int hband = 1 << 30;
    KB_CUDA(ctx, cudaMemcpyAsync(&hband, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *band_out = hband <= 64 ? ((hband + 15) / 16) * 16 : -1;
    return KB_OK;
}

int sample_plan_create(kb_ctx* ctx, Scope& sc, int p, int m, const double* dmu, const double* dSiS, const double* dLb,
                       long ldo, long chunk, bool need_zbuf, SamplePlan* plan) {
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);   // modelX.jl:105
    plan->p = p; plan->m = m;
    plan->chunk = (chunk + 1) & ~1L;
    plan->ld = (p + 1) & ~1L;
    plan->dLb = dLb; plan->ldo = ldo;
    KB_TRY(sc.alloc(&plan->nSiS, (size_t)plan->ld * p));
    KB_TRY(sc.alloc(&plan->bias, p));
    if (m > 1) {
        // Structure of the factor blocks: for Diagonal(s) every M_k = L_kk − S inv(L_kk)' has the strict lower triangle of
        // L_kk (bit for bit), i.e. M_k = L_kk + N_k with N_k upper triangular.  Then, with S_k = Σ_{i>=k} Z_i,
        //     Z_k L_kk + (Σ_{i>k} Z_i) M_k = S_k L_kk + S_{k+1} N_k
        // is TWO triangular products (2 n p² flops per copy) instead of a triangular and a dense one (3 n p²).
        // Detected here, so kb_sample also takes the short form for factor blocks passed in by the caller;
        // anything else (dense S of the group path) keeps the general form.
        static const bool allow = [] { const char* e = getenv("KB_SAMPLE_STRUCTURED"); return !(e && e[0] == '0'); }();
        const size_t bstride = (size_t)ldo * p;
        // group blocks (contiguous groups of size g): M_k = L_kk − S inv(L_kk)' differs from L_kk at most g − 1 rows
        // below the diagonal; a band of up to 64 rows still leaves the second product (almost) triangular
        int band = -1;
        KB_TRY(factor_band_dev(ctx, dLb, ldo, p, m, &band));
        plan->structured = allow && band >= 0;
        plan->band = plan->structured ? band : 0;
        KB_TRY(sc.alloc(&plan->Mu, (size_t)plan->chunk * p));
        KB_TRY(sc.alloc(&plan->T, (size_t)plan->chunk * p * (plan->structured ? m : m - 1)));
        if (plan->structured) {
            KB_TRY(sc.alloc(&plan->N, (size_t)plan->ld * p * (m - 1)));
            factor_upper_part_kernel<<<ew_grid2(ctx, (long)p * p * (m - 1)), 256, 0, ctx->stream>>>(dLb, ldo, (long)bstride, p, m,
                                                                                                  plan->band, plan->N, plan->ld);
            ctx->launches++;
            // operand tiles of the batched launch staged by TMA (cp.async.bulk.tensor + mbarrier ring) when the buffers
            // qualify (16-byte aligned, even strides) and KB_SAMPLE_TMA=1 asks for it (default: the LDGSTS kernel, which measured
            // 0.6 % faster in this stage)
            const long tstride = (long)plan->chunk * p;
            if (sample_tma_usable(plan->T, plan->chunk, tstride, dLb, ldo, 2 * (long)bstride, plan->N, plan->ld, plan->ld * (long)p) &&
                sample_tma_maps(ctx, &plan->maps, plan->T, plan->chunk, p, m, tstride, dLb, ldo, (long)bstride, plan->N, plan->ld) == KB_OK)
                plan->tma = true;
        }
    }
    if (need_zbuf) KB_TRY(sc.alloc(&plan->Zbuf, (size_t)plan->chunk * p * m));
    negate_kernel<<<ew_grid2(ctx, (long)p * p), 256, 0, ctx->stream>>>(dSiS, ldo, p, p, plan->nSiS, plan->ld);
    ctx->launches++;
    mu_sis_kernel<<<p, 256, 0, ctx->stream>>>(dmu, dSiS, ldo, p, plan->bias);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int sample_chunk(kb_ctx* ctx, const SamplePlan& pl, const double* Xc, long rows, long ldx, const double* Zc, long ldzc,
                 uint64_t seed, int64_t row_global0, double* Xko, long ldxko) {
    if (rows <= 0) return KB_OK;
    if (rows > pl.chunk) return set_error(ctx, KB_ERR_INVALID, "sample_chunk: %ld rows exceed the plan's chunk", rows);
    const int p = pl.p, m = pl.m;
    const size_t bstride = (size_t)pl.ldo * p;
    if (!Zc) {
        if (!pl.Zbuf) return set_error(ctx, KB_ERR_INVALID, "sample_chunk: plan has no Z buffer");
        KB_TRY(randn_dev(ctx, seed, row_global0, rows, (long)p * m, 0, pl.Zbuf, pl.chunk));
        Zc = pl.Zbuf; ldzc = pl.chunk;
    }
    if (m == 1) {
        // X̃ = X − X ΣinvS + μ'ΣinvS + Z L        (modelX.jl:112) : ONE fused two-segment GEMM
        GemmParams P{};
        P.M = (int)rows; P.N = p; P.nseg = 2;
        P.seg[0] = GemmSeg{Xc, ldx, pl.nSiS, pl.ld, p, 0};
        P.seg[1] = GemmSeg{Zc, ldzc, pl.dLb, pl.ldo, p, SEG_KLO_N};     // L lower: L(k,n) = 0 for k < n
        P.C = Xko; P.ldc = ldxko; P.Cin = Xc; P.ldcin = ldx; P.alpha = 1.0; P.beta = 1.0; P.bias_n = pl.bias;
        return gemm(ctx, A_MCONTIG, B_KCONTIG, P);
    }
    // μi = X − (X .− μ')ΣinvS once per row chunk (modelX.jl:118)
    {
        GemmParams P{};
        P.M = (int)rows; P.N = p; P.nseg = 1;
        P.seg[0] = GemmSeg{Xc, ldx, pl.nSiS, pl.ld, p, 0};
        P.C = pl.Mu; P.ldc = pl.chunk; P.Cin = Xc; P.ldcin = ldx; P.alpha = 1.0; P.beta = 1.0; P.bias_n = pl.bias;
        KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, P));
    }
    // copy k (columns k p .. (k+1) p of X̃): μi + Z_k L_kk + (Σ_{i>k} Z_i) M_k  (block column k of
    // randn(n, mp) * L, modelX.jl:121).  The suffix sums T_k = Σ_{i>k} Z_i are formed first, then ALL m copies run as
    // ONE batched launch (gridDim.z = m: 5 x 5056 CTAs at p = 5000 instead of 5 launches with a partial last wave
    // each); the last copy has no M block (one segment).
    const size_t tstride = (size_t)pl.chunk * p;
    if (pl.structured) {
        // X̃_k = μi + S_k L_kk + S_{k+1} N_k  (S_k inclusive suffix sums; L_kk lower, N_k upper: every tile's two clipped
        // k-ranges add up to p + BN, so the batched launch is perfectly balanced)
        suffix_sum_incl_kernel<<<ew_grid2(ctx, rows * p), 256, 0, ctx->stream>>>(pl.T, pl.chunk, (long)tstride, Zc, ldzc,
                                                                                 (long)p * ldzc, rows, p, m);
        ctx->launches++;
        if (pl.tma) return sample_tma_launch(ctx, pl.maps, (int)rows, p, m, pl.band, pl.Mu, pl.chunk, Xko, ldxko);
        GemmParams P{};
        P.M = (int)rows; P.N = p;
        P.seg[0] = GemmSeg{pl.T, pl.chunk, pl.dLb, pl.ldo, p, SEG_KLO_N};                       // S_k L_kk
        P.seg[1] = GemmSeg{pl.T + tstride, pl.chunk, pl.N, pl.ld, p, SEG_KHI_N};                // S_{k+1} N_k
        P.nseg = 2;
        P.batch = m; P.last_nseg = 1; P.khi_slack = pl.band;
        P.zA[0] = (long)tstride; P.zB[0] = 2 * (long)bstride;
        P.zA[1] = (long)tstride; P.zB[1] = (long)pl.ld * p;
        P.zC = (long)p * ldxko;
        P.C = Xko; P.ldc = ldxko; P.Cin = pl.Mu; P.ldcin = pl.chunk;
        P.alpha = 1.0; P.beta = 1.0; P.bias_n = nullptr;
        KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, P));
        KB_CUDA(ctx, cudaGetLastError());
        return KB_OK;
    }
    suffix_sum_kernel<<<ew_grid2(ctx, rows * p), 256, 0, ctx->stream>>>(pl.T, pl.chunk, (long)tstride, Zc, ldzc, (long)p * ldzc,
                                                                        rows, p, m);
    ctx->launches++;
    {
        GemmParams P{};
        P.M = (int)rows; P.N = p;
        P.seg[0] = GemmSeg{Zc, ldzc, pl.dLb, pl.ldo, p, SEG_KLO_N};            // Z_k L_kk, L lower
        P.seg[1] = GemmSeg{pl.T, pl.chunk, pl.dLb + bstride, pl.ldo, p, 0};    // T_k M_k
        P.nseg = 2;
        P.batch = m; P.last_nseg = 1;
        P.zA[0] = (long)p * ldzc; P.zB[0] = 2 * (long)bstride;
        P.zA[1] = (long)tstride;  P.zB[1] = 2 * (long)bstride;
        P.zC = (long)p * ldxko;
        P.C = Xko; P.ldc = ldxko; P.Cin = pl.Mu; P.ldcin = pl.chunk;
        P.alpha = 1.0; P.beta = 1.0; P.bias_n = nullptr;
        KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, P));
    }
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int sample_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, const double* dmu, const double* dSiS,
               const double* dLb, long ldo, int m, const double* dZ, long ldz, uint64_t seed, int64_t row_offset,
               double* dXko, long ldxko) {
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    if (n <= 0) return KB_OK;
    long chunk = ctx->sample_chunk_rows > 0 ? ctx->sample_chunk_rows : 4096;
    if (chunk > n) chunk = n;
    Scope sc(ctx);
    SamplePlan plan;
    KB_TRY(sample_plan_create(ctx, sc, p, m, dmu, dSiS, dLb, ldo, chunk, dZ == nullptr, &plan));
    for (long r0 = 0; r0 < n; r0 += plan.chunk) {
        const long rows = std::min(plan.chunk, n - r0);
        KB_TRY(sample_chunk(ctx, plan, dX + r0, rows, ldx, dZ ? dZ + r0 : nullptr, ldz, seed, row_offset + r0,
                            dXko + r0, ldxko));
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // scoped buffers are freed on return
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------
// Gram
// ------------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const double* X, long n, long ldx, int p, double* colsum, int accumulate) {
    const int c = blockIdx.x;
    __shared__ double red[256];
    const double* x = X + (long)c * ldx;
    double a0 = 0.0, a1 = 0.0;
    long i = threadIdx.x;
    for (; i + 256 < n; i += 512) { a0 += x[i]; a1 += x[i + 256]; }
    if (i < n) a0 += x[i];
    red[threadIdx.x] = a0 + a1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) colsum[c] = accumulate ? colsum[c] + red[0] : red[0];
}

int gram_partial_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, double* dG, long ldg, double* dcolsum,
                     int accumulate) {
    if (n <= 0) return KB_OK;
    // lower tiles of G (+)= X' X :  A(m,k) = X(k,m), B(k,n) = X(k,n), K = local rows
    // K is split into segments of < 2^31 rows implicitly: n fits int for every supported shard.
    if (n > 2000000000L) return set_error(ctx, KB_ERR_INVALID, "gram: more than 2e9 rows per call");
    KB_TRY(gemm1(ctx, A_KCONTIG, B_KCONTIG, p, p, (int)n, 1.0, dX, ldx, dX, ldx, accumulate ? 1.0 : 0.0, dG, ldg, 0, 1, 0));
    if (dcolsum) {
        colsum_kernel<<<p, 256, 0, ctx->stream>>>(dX, n, ldx, p, dcolsum, accumulate);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
    }
    return KB_OK;
}

// G (lower valid) -> (G − colsum colsum'/n_total) / denom on the lower triangle, mirrored to the upper;
// colsum -> column means.
__global__ void gram_finish_kernel(double* G, long ldg, const double* colsum, double n_total, int p, int center,
                                   double denom) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        if (i < k) continue;
        double v = G[i + (long)k * ldg];
        if (center) v -= colsum[i] * colsum[k] / n_total;
        G[i + (long)k * ldg] = v / denom;
    }
}
__global__ void scale_vec_kernel(double* x, int p, double a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p) x[i] *= a;
}

int gram_finish_dev(kb_ctx* ctx, double* dG, long ldg, double* dcolsum, long n_total, int p, int center, double denom) {
    if (!(denom > 0.0)) denom = (double)n_total;
    if (center && !dcolsum) return set_error(ctx, KB_ERR_INVALID, "gram_finish: centring needs the column sums");
    gram_finish_kernel<<<ew_grid2(ctx, (long)p * p), 256, 0, ctx->stream>>>(dG, ldg, dcolsum, (double)n_total, p, center, denom);
    ctx->launches++;
    KB_TRY(mirror_lower_to_upper(ctx, dG, ldg, p));
    if (dcolsum) {
        scale_vec_kernel<<<ceil_div(p, 256), 256, 0, ctx->stream>>>(dcolsum, p, 1.0 / (double)n_total);
        ctx->launches++;
    }
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------
// Ledoit-Wolf linear shrinkage toward diag(S)  (CovarianceEstimation.jl LinearShrinkage(DiagonalUnequalVariance(), :lw),
// the default covariance_approximator of the 2-argument entry, modelX.jl:43-50).  Restated from the published
// algorithm (the package is not under /root/reference: parity unpinned):
//   S = Xc'Xc / n,  pi_ij = (1/n) sum_k (x_ki x_kj)^2 − S_ij^2,  lambda = clamp( (sum_{i≠j} pi_ij / n) / sum_{i≠j} S_ij^2, 0, 1 ),
//   Sigma = (1 − lambda) S + lambda diag(S).
// Both Grams (of Xc and of Xc∘Xc) run on the DMMA GEMM and shard by rows like kb_gram.
// ------------------------------------------------------------------------------------------------
// Y = (X − mean)^2 for a row slab
__global__ void centered_square_kernel(const double* X, long n, long ldx, int p, const double* mean, double* Y, long ldy) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = n * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % n, c = idx / n;
        const double d = X[i + c * ldx] - mean[c];
        Y[i + c * ldy] = d * d;
    }
}
int gram_sq_partial_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, const double* dmean, double* dY, long ldy,
                        double* dG2, long ldg, int accumulate) {
    if (n <= 0) return KB_OK;
    centered_square_kernel<<<ew_grid2(ctx, n * p), 256, 0, ctx->stream>>>(dX, n, ldx, p, dmean, dY, ldy);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return gemm1(ctx, A_KCONTIG, B_KCONTIG, p, p, (int)n, 1.0, dY, ldy, dY, ldy, accumulate ? 1.0 : 0.0, dG2, ldg, 0, 1, 0);
}
// sums over the strict lower triangle: out[0] += sum (G2_ij / n − S_ij^2), out[1] += sum S_ij^2
__global__ void lw_sums_kernel(const double* S, long lds, const double* G2, long ldg, int p, double n, double* out) {
    __shared__ double r0[256], r1[256];
    double a = 0.0, b = 0.0;
    const long total = (long)p * p;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        if (i > k) {
            const double s = S[i + (long)k * lds];
            a += G2[i + (long)k * ldg] / n - s * s;
            b += s * s;
        }
    }
    r0[threadIdx.x] = a; r1[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = r0[0]; out[2 * blockIdx.x + 1] = r1[0]; }
}
__global__ void lw_shrink_kernel(double* S, long lds, int p, double lambda) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), k = (int)(idx / p);
        if (i != k) S[i + (long)k * lds] *= (1.0 - lambda);
    }
}
// dS: finished covariance (full symmetric, kb_gram_finish_dev output); dG2: raw sums of the squared-centred Gram
// (lower triangle valid).  Writes the shrunk covariance into dS and returns lambda.
int lw_finish_dev(kb_ctx* ctx, double* dS, long lds, const double* dG2, long ldg, int p, long n_total, double* lambda_out) {
    Scope sc(ctx);
    const int nb = 256;
    double* dout;
    KB_TRY(sc.alloc(&dout, 2 * nb));
    lw_sums_kernel<<<nb, 256, 0, ctx->stream>>>(dS, lds, dG2, ldg, p, (double)n_total, dout);
    ctx->launches++;
    std::vector<double> h(2 * nb);
    KB_CUDA(ctx, cudaMemcpyAsync(h.data(), dout, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double num = 0.0, den = 0.0;
    for (int i = 0; i < nb; ++i) { num += h[2 * i]; den += h[2 * i + 1]; }
    double lam = (den > 0.0) ? (num / (double)n_total) / den : 0.0;     // the factor 2 of (i,j)/(j,i) cancels
    lam = lam < 0.0 ? 0.0 : (lam > 1.0 ? 1.0 : lam);
    lw_shrink_kernel<<<ew_grid2(ctx, (long)p * p), 256, 0, ctx->stream>>>(dS, lds, p, lam);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    *lambda_out = lam;
    return KB_OK;
}

int sym_upper_to_full_dev(kb_ctx* ctx, const double* dA, long lda, int p, double* dOut, long ldo) {
    sym_upper_to_full_kernel<<<ew_grid2(ctx, (long)p * p), 256, 0, ctx->stream>>>(dA, lda, p, dOut, ldo);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace kb
