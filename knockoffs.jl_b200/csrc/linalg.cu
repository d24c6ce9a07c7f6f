// Dense FP64 building blocks: GEMM launcher, blocked Cholesky, triangular inverse.
#include "linalg.cuh"
#include <cstdarg>

namespace kb {

int set_error(kb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}

// ------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------
template <int AL, int BL, bool VEC>
static int launch_gemm_t(kb_ctx* ctx, const GemmParams& P) {
    static bool configured = false;
    auto kern = dgemm_dmma_kernel<AL, BL, VEC>;
    if (!configured) {
        KB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
        configured = true;
    }
    dim3 grid(ceil_div(P.M, GEMM_BM), ceil_div(P.N, GEMM_BN));
    kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(P);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int gemm(kb_ctx* ctx, int AL, int BL, const GemmParams& P) {
    if (P.M <= 0 || P.N <= 0) return KB_OK;
    bool vec = true;
    for (int s = 0; s < P.nseg; ++s) {
        const GemmSeg& g = P.seg[s];
        if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda & 1) || (g.ldb & 1)) vec = false;
    }
#define KB_DISPATCH(al, bl)                                         \
    if (AL == al && BL == bl)                                       \
        return vec ? launch_gemm_t<al, bl, true>(ctx, P) : launch_gemm_t<al, bl, false>(ctx, P);
    KB_DISPATCH(A_MCONTIG, B_KCONTIG)
    KB_DISPATCH(A_MCONTIG, B_NCONTIG)
    KB_DISPATCH(A_KCONTIG, B_KCONTIG)
    KB_DISPATCH(A_KCONTIG, B_NCONTIG)
#undef KB_DISPATCH
    return set_error(ctx, KB_ERR_INVALID, "gemm: bad layout");
}

int gemm1(kb_ctx* ctx, int AL, int BL, int M, int N, int K, double alpha, const double* A, long lda,
          const double* B, long ldb, double beta, double* C, long ldc, int seg_flags, int tile_mode,
          int store_mode) {
    GemmParams P{};
    P.M = M; P.N = N; P.nseg = 1;
    P.seg[0] = GemmSeg{A, lda, B, ldb, K, seg_flags};
    P.C = C; P.ldc = ldc; P.Cin = C; P.ldcin = ldc;
    P.alpha = alpha; P.beta = beta; P.bias_n = nullptr;
    P.tile_mode = tile_mode; P.store_mode = store_mode;
    return gemm(ctx, AL, BL, P);
}

// ------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------
__global__ void zero_strict_upper_kernel(double* A, long ld, int p) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int i = (int)(idx % p), k = (int)(idx / p);
        if (i < k) A[i + (long)k * ld] = 0.0;
    }
}
__global__ void mirror_lower_kernel(double* A, long ld, int p) {
    // tiled transpose-copy of the strict lower triangle into the upper
    __shared__ double t[32][33];
    int bi = blockIdx.x, bk = blockIdx.y;
    if (bi < bk) return;
    int i = bi * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = bk * 32 + r;
        t[r][threadIdx.x] = (i < p && k < p) ? A[i + (long)k * ld] : 0.0;   // (row i, col k)
    }
    __syncthreads();
    int k2 = bk * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i2 = bi * 32 + r;
        if (i2 < p && k2 < p && k2 < i2) A[k2 + (long)i2 * ld] = t[threadIdx.x][r];
    }
}
__global__ void copy_matrix_kernel(const double* A, long lda, double* B, long ldb, int rows, int cols) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int i = (int)(idx % rows), k = (int)(idx / rows);
        B[i + (long)k * ldb] = A[i + (long)k * lda];
    }
}

static inline int ew_grid(kb_ctx* ctx, long total, int threads = 256) {
    long g = (total + threads - 1) / threads;
    long cap = (long)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int zero_strict_upper(kb_ctx* ctx, double* A, long ld, int p) {
    zero_strict_upper_kernel<<<ew_grid(ctx, (long)p * p), 256, 0, ctx->stream>>>(A, ld, p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}
int mirror_lower_to_upper(kb_ctx* ctx, double* A, long ld, int p) {
    dim3 grid(ceil_div(p, 32), ceil_div(p, 32)), block(32, 8);
    mirror_lower_kernel<<<grid, block, 0, ctx->stream>>>(A, ld, p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}
int copy_matrix(kb_ctx* ctx, const double* A, long lda, double* B, long ldb, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return KB_OK;
    copy_matrix_kernel<<<ew_grid(ctx, (long)rows * cols), 256, 0, ctx->stream>>>(A, lda, B, ldb, rows, cols);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}
int fill_zero(kb_ctx* ctx, double* A, size_t count) {
    KB_CUDA(ctx, cudaMemsetAsync(A, 0, count * sizeof(double), ctx->stream));
    return KB_OK;
}

// ------------------------------------------------------------------------------------------
// Cholesky: diagonal block kernel (one CTA): factor nb x nb block in shared memory, write L back,
// then invert it in place and write inv(L_bb) to `inv` (ld = POTRF_NB).
// ------------------------------------------------------------------------------------------
constexpr int DIAG_PITCH = POTRF_NB + 1;
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_PANEL = 16;     // columns factored per panel step
constexpr size_t DIAG_SMEM = (size_t)(POTRF_NB * DIAG_PITCH + POTRF_NB) * sizeof(double);

// One CTA factors an nb x nb (<= 128) diagonal block held in shared memory, writes L back and then
// inv(L) to `inv`.  Blocked right-looking inside the CTA: a 16-column panel is factored by its own
// 128 x 16 thread strip (one thread per row, columns swept with CTA barriers), then the trailing
// matrix gets ONE rank-16 update with 4 x 4 register tiles (no integer division in any loop).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* A, long ld, int nb, double* inv, int* fail, int base) {
    extern __shared__ double dsm[];
    double* T = dsm;                              // T(i,k) at k*DIAG_PITCH + i
    double* xv = dsm + POTRF_NB * DIAG_PITCH;
    const int tid = threadIdx.x;
#define T_(i, k) T[(k) * DIAG_PITCH + (i)]
    {
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB)
            T_(i, k) = (i < nb && k < nb) ? ((i >= k) ? A[i + (long)k * ld] : 0.0) : (i == k ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j0 = 0; j0 < nb; j0 += DIAG_PANEL) {
        const int jb = min(DIAG_PANEL, nb - j0);
        // ---- panel: columns j0 .. j0+jb-1, rows j0 .. nb-1; thread `tid` < 128 owns row tid ----
        for (int j = j0; j < j0 + jb; ++j) {
            double d = T_(j, j);
            if (!(d > 0.0)) {                      // also catches NaN
                if (tid == 0) atomicCAS(fail, 0, base + j + 1);
                d = 1.0;
            }
            const double ljj = sqrt(d);
            __syncthreads();
            if (tid < nb && tid >= j) {
                const double v = (tid == j) ? ljj : T_(tid, j) / ljj;
                T_(tid, j) = v;
            }
            __syncthreads();
            // update the remaining panel columns k = j+1 .. j0+jb-1 of rows >= k
            if (tid < nb && tid > j) {
                const double lij = T_(tid, j);
                for (int k = j + 1; k < j0 + jb; ++k)
                    if (tid >= k) T_(tid, k) -= lij * T_(k, j);
            }
            __syncthreads();
        }
        // ---- trailing update: T(i,k) -= sum_{c in panel} T(i,c) T(k,c) for i >= k >= j0+jb ----
        const int t0 = j0 + jb;
        if (t0 < nb) {
            // 4x4 tiles over the trailing square; 32 x 32 tile grid covers 128 x 128, 512 threads -> 2 tiles each
            for (int tile = tid; tile < 32 * 32; tile += DIAG_THREADS) {
                const int ti = t0 + (tile & 31) * 4, tk = t0 + (tile >> 5) * 4;
                if (ti >= nb || tk >= nb || ti + 3 < tk) continue;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
                for (int c = j0; c < t0; ++c) {
                    double li[4], lk[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { li[a] = T_(ti + a, c); lk[a] = T_(tk + a, c); }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(li[a], lk[b], acc[a][b]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (ti + a < nb && tk + b < nb && ti + a >= tk + b) T_(ti + a, tk + b) -= acc[a][b];
            }
        }
        __syncthreads();
    }
    {
        const int i = tid & (POTRF_NB - 1);
        if (i < nb)
            for (int k = tid >> 7; k < nb; k += DIAG_THREADS / POTRF_NB)
                A[i + (long)k * ld] = T_(i, k);    // strict upper of the block becomes 0
    }
    __syncthreads();
    // in-place inverse of lower-triangular T (dtrti2, columns from the right)
    const int part = tid & 3, row = tid >> 2;      // 4 threads per row
    for (int j = nb - 1; j >= 0; --j) {
        const double ljj = T_(j, j);
        for (int i = j + 1 + tid; i < nb; i += DIAG_THREADS) xv[i] = T_(i, j);
        __syncthreads();
        double acc = 0.0;
        if (row > j && row < nb)
            for (int k = j + 1 + part; k <= row; k += 4) acc += T_(row, k) * xv[k];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0) {
            if (row > j && row < nb) T_(row, j) = -acc / ljj;
            else if (row == j) T_(j, j) = 1.0 / ljj;
        }
        __syncthreads();
    }
    {
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB)
            inv[i + (long)k * POTRF_NB] = (i < nb && k < nb && i >= k) ? T_(i, k) : 0.0;
    }
#undef T_
}

int potrf_lower(kb_ctx* ctx, double* A, long ld, int p, double* invdiag, int* d_fail) {
    static bool configured = false;
    if (!configured) {
        KB_CUDA(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)DIAG_SMEM));
        configured = true;
    }
    int blk = 0;
    for (int r0 = 0; r0 < p; r0 += POTRF_NB, ++blk) {
        const int nb = (p - r0 < POTRF_NB) ? p - r0 : POTRF_NB;
        double* Abb = A + r0 + (long)r0 * ld;
        double* inv = invdiag + (size_t)blk * POTRF_NB * POTRF_NB;
        potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(Abb, ld, nb, inv, d_fail, r0);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
        const int rem = p - r0 - nb;
        if (rem <= 0) break;
        double* A21 = A + (r0 + nb) + (long)r0 * ld;
        // L21 = A21 * inv(L11)'   (in place: one N tile, each CTA reads only the rows it writes)
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, nb, nb, 1.0, A21, ld, inv, POTRF_NB, 0.0, A21, ld));
        // A22 -= L21 L21'  (lower tiles)
        double* A22 = A + (r0 + nb) + (long)(r0 + nb) * ld;
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, rem, nb, -1.0, A21, ld, A21, ld, 1.0, A22, ld, 0, 1, 0));
    }
    return zero_strict_upper(ctx, A, ld, p);
}

// ------------------------------------------------------------------------------------------
// triangular inverse by recursive doubling on block pairs
// ------------------------------------------------------------------------------------------
__global__ void place_diag_inverses_kernel(const double* invdiag, double* W, long ldw, int p) {
    const int blk = blockIdx.x;
    const int r0 = blk * POTRF_NB;
    const int nb = min(POTRF_NB, p - r0);
    const double* inv = invdiag + (size_t)blk * POTRF_NB * POTRF_NB;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        int i = e % nb, k = e / nb;
        W[(r0 + i) + (long)(r0 + k) * ldw] = inv[i + (long)k * POTRF_NB];
    }
}

int trtri_lower(kb_ctx* ctx, const double* L, long ldl, int p, const double* invdiag, double* W, long ldw,
                double* work) {
    KB_TRY(fill_zero(ctx, W, (size_t)ldw * p));
    place_diag_inverses_kernel<<<ceil_div(p, POTRF_NB), 256, 0, ctx->stream>>>(invdiag, W, ldw, p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    for (long b = POTRF_NB; b < p; b *= 2) {
        for (long o = 0; o + b < p; o += 2 * b) {
            const int b1 = (int)b;
            const int b2 = (int)((p - o - b < b) ? p - o - b : b);
            const long ldt = (b2 + 1) & ~1L;
            const double* L21 = L + (o + b1) + o * ldl;
            const double* W11 = W + o + o * ldw;
            const double* W22 = W + (o + b1) + (o + b1) * ldw;
            double* W21 = W + (o + b1) + o * ldw;
            // T = L21 * W11          (W11 lower: B(k,n) = 0 for k < n)
            KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, b2, b1, b1, 1.0, L21, ldl, W11, ldw, 0.0, work, ldt, SEG_KLO_N));
            // W21 = -W22 * T         (W22 lower: A(m,k) = 0 for k > m)
            KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, b2, b1, b2, -1.0, W22, ldw, work, ldt, 0.0, W21, ldw, SEG_KHI_M));
        }
    }
    return KB_OK;
}

int syrk_WtW_lower(kb_ctx* ctx, const double* W, long ldw, int p, double* Out, long ldo) {
    // Out(m,n) = sum_k W(k,m) W(k,n),  W lower  =>  k >= max(m,n); lower tiles only.
    return gemm1(ctx, A_KCONTIG, B_KCONTIG, p, p, p, 1.0, W, ldw, W, ldw, 0.0, Out, ldo, SEG_KLO_M | SEG_KLO_N, 1, 0);
}

}  // namespace kb
