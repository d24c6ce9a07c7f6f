// Dense FP64 building blocks: GEMM launcher, blocked Cholesky, triangular inverse.
#include "linalg.cuh"
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include <map>
#include <mutex>

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

cudaError_t raise_max_dynamic_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> current;     // the attribute is per DEVICE and per function
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = current[std::make_pair(dev, func)];
    if (bytes <= cur || bytes <= 48 * 1024) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

// ------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------
// CTA configuration (build-time; measured on B200 at p = 5000, profiles/r01_gemm_variants.md):
//   default 64 x 64 / 4 warps / 3 CTAs per SM;  KB_GEMM_CFG=128: 128 x 128 / 8 warps / 1 CTA;  =6412: 64 x 128 / 4 warps / 2 CTAs
#if defined(KB_GEMM_CFG) && KB_GEMM_CFG == 128
using GemmDefaultCfg = GemmCfg<128, 128, 2, 4, 1>;
#elif defined(KB_GEMM_CFG) && KB_GEMM_CFG == 6412
using GemmDefaultCfg = GemmCfg<64, 128, 2, 2, 2>;
#else
#ifndef KB_GEMM_MINCTAS
#define KB_GEMM_MINCTAS 3
#endif
using GemmDefaultCfg = GemmCfg<64, 64, 2, 2, KB_GEMM_MINCTAS>;
#endif

template <int AL, int BL, bool VEC>
static int launch_gemm_t(kb_ctx* ctx, const GemmParams& P) {
    using Cfg = GemmDefaultCfg;
    auto kern = dgemm_dmma_kernel<Cfg, AL, BL, VEC>;
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)kern, Cfg::SMEM_BYTES));
    dim3 grid(ceil_div(P.M, Cfg::BM), ceil_div(P.N, Cfg::BN), P.batch > 1 ? P.batch : 1);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(P);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

int gemm(kb_ctx* ctx, int AL, int BL, const GemmParams& P) {
    if (P.M <= 0 || P.N <= 0) return KB_OK;
    bool vec = true;
    for (int s = 0; s < P.nseg; ++s) {
        const GemmSeg& g = P.seg[s];
        // executed flops: a triangular (k-clipped) operand or a lower-tiles-only output halves the work
        double f = 2.0 * (double)P.M * (double)P.N * (double)g.K;
        if (g.flags) f *= 0.5;
        if (P.tile_mode == 1) f *= 0.5;
        const int nb = P.batch > 1 ? P.batch : 1;
        ctx->flops += f * ((P.batch > 1 && P.last_nseg > 0 && s >= P.last_nseg) ? nb - 1 : nb);
        if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda & 1) || (g.ldb & 1)) vec = false;
        if (P.batch > 1 && ((P.zA[s] & 1) || (P.zB[s] & 1))) vec = false;
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
constexpr int DIAG_PITCH = POTRF_NB + 4;   // pitch mod 16 == 4: DMMA fragment loads are bank-conflict free
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_PANEL = 16;     // columns factored per panel step
constexpr int DIAG_NPANEL = POTRF_NB / DIAG_PANEL;
// T (128 x 129) + W (inverse, lower triangle folded into a 129 x 64 rectangle) + 16 x 17 scratch for the panel's
// diagonal block and its inverse: 132 + 66 + 4 KB
constexpr size_t DIAG_SMEM = (size_t)(POTRF_NB * DIAG_PITCH + DIAG_PITCH * (POTRF_NB / 2) + 2 * DIAG_PANEL * (DIAG_PANEL + 1)) * sizeof(double);
// folded lower-triangular storage: (i, k), i >= k.  Columns k < 64 keep rows k..127 at row i + 1; columns k >= 64 go
// to column 127 - k, row 127 - i (the part of that column the first half leaves free).
__device__ __forceinline__ int wfold(int i, int k) {
    return (k < POTRF_NB / 2) ? (k * DIAG_PITCH + i + 1) : ((POTRF_NB - 1 - k) * DIAG_PITCH + (POTRF_NB - 1 - i));
}

// C(i0.., j0..) (+)= sign * A(i0.., k-range) * B(j0.., k-range)'  over a rows x cols output block, 4 x 4 register
// tiles dealt round-robin to the CTA's threads.  All operands in shared memory with pitch DIAG_PITCH (column-major),
// element (i, k) of X at X[k * DIAG_PITCH + i].  lower_only: skip tiles strictly above the diagonal of the output.
template <bool ACCUM, bool LOWER_ONLY>
__device__ __forceinline__ void smem_tile_gemm_nt(double* C, int ci0, int cj0, int rows, int cols, const double* A, int ai0,
                                                  int ak0, const double* B, int bj0, int bk0, int kk, double sign, int tid) {
    const int tr = (rows + 3) >> 2, tc = (cols + 3) >> 2;
    for (int tile = tid; tile < tr * tc; tile += DIAG_THREADS) {
        const int ti = (tile % tr) * 4, tj = (tile / tr) * 4;
        if (LOWER_ONLY && ci0 + ti + 3 < cj0 + tj) continue;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int k = 0; k < kk; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                av[a] = (ti + a < rows) ? A[(ak0 + k) * DIAG_PITCH + ai0 + ti + a] : 0.0;
                bv[a] = (tj + a < cols) ? B[(bk0 + k) * DIAG_PITCH + bj0 + tj + a] : 0.0;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (ti + a < rows && tj + b < cols && (!LOWER_ONLY || ci0 + ti + a >= cj0 + tj + b)) {
                    double* c = &C[(cj0 + tj + b) * DIAG_PITCH + ci0 + ti + a];
                    *c = ACCUM ? (*c + sign * acc[a][b]) : sign * acc[a][b];
                }
            }
    }
}

// One CTA factors an nb x nb (<= 128) diagonal block held in shared memory, writes L back and then inv(L) to `inv`.
// Blocked right-looking inside the CTA with 16-column panels:
//   (a) warp 0 factors the 16 x 16 diagonal block and inverts it (warp-synchronous, no CTA barrier per column),
//   (b) the panel below it is multiplied by that inverse (TRSM as a small GEMM),
//   (c) the trailing matrix gets one rank-16 update with 4 x 4 register tiles.
// The triangular inverse is blocked the same way: 16 x 16 diagonal inverses from (a), then recursive doubling
// W21 = -W22 (L21 W11) at block sizes 16, 32, 64.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel_v1(double* A, long ld, int nb, double* inv, int* fail, int base) {
    extern __shared__ double dsm[];
    double* T = dsm;                                   // T(i,k) at k*DIAG_PITCH + i
    double* W = dsm + POTRF_NB * DIAG_PITCH;           // inverse being assembled (lower)
    double* D16 = W + DIAG_PITCH * (POTRF_NB / 2);     // 16 x 17: factored diagonal block of the current panel
    double* I16 = D16 + DIAG_PANEL * (DIAG_PANEL + 1); // 16 x 17: its inverse
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#define T_(i, k) T[(k) * DIAG_PITCH + (i)]
#define W_(i, k) W[wfold((i), (k))]
#define WG_(i, k) (((i) >= (k)) ? W[wfold((i), (k))] : 0.0)
#define D_(i, k) D16[(k) * (DIAG_PANEL + 1) + (i)]
#define I_(i, k) I16[(k) * (DIAG_PANEL + 1) + (i)]
    {
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB) {
            T_(i, k) = (i < nb && k < nb) ? ((i >= k) ? A[i + (long)k * ld] : 0.0) : (i == k ? 1.0 : 0.0);
            if (i >= k) W_(i, k) = 0.0;
        }
    }
    __syncthreads();
    for (int jb = 0; jb < DIAG_NPANEL; ++jb) {
        const int j0 = jb * DIAG_PANEL;
        // ---- (a) warp 0: 16 x 16 Cholesky + inverse, lane r owns row r (lanes 16..31 idle) ----
        if (warp == 0) {
            const int r = lane;
            double row[DIAG_PANEL], dinv[DIAG_PANEL];
#pragma unroll
            for (int c = 0; c < DIAG_PANEL; ++c) row[c] = (r < DIAG_PANEL && c <= r) ? T_(j0 + r, j0 + c) : 0.0;
#pragma unroll
            for (int c = 0; c < DIAG_PANEL; ++c) {
                double d = __shfl_sync(0xffffffffu, row[c], c);          // pivot (row c, column c)
                if (!(d > 0.0)) {                                         // also catches NaN
                    if (lane == 0 && j0 + c < nb) atomicCAS(fail, 0, base + j0 + c + 1);
                    d = 1.0;
                }
                // 1/sqrt(d) once (no FP64 divide, no FP64 sqrt sequence on the serial chain); L_cc = d * rsqrt(d)
                const double rinv = rsqrt(d);
                dinv[c] = rinv;
                row[c] = (r == c) ? d * rinv : row[c] * rinv;             // column c scaled (rows > c meaningful)
#pragma unroll
                for (int k = c + 1; k < DIAG_PANEL; ++k) {
                    const double lkc = __shfl_sync(0xffffffffu, row[c], k);   // L(k, c)
                    if (r >= k) row[k] = fma(-row[c], lkc, row[k]);
                }
            }
            if (r < DIAG_PANEL) {
#pragma unroll
                for (int c = 0; c < DIAG_PANEL; ++c) {
                    const double lv = (c <= r) ? row[c] : 0.0;
                    T_(j0 + r, j0 + c) = lv;
                    D_(r, c) = lv;
                }
            }
            __syncwarp();
            // inverse: lane c computes COLUMN c of inv(L) by forward substitution (L read from shared memory,
            // all lanes read the same address -> broadcast)
            if (r < DIAG_PANEL) {
                const int c = r;
                double x[DIAG_PANEL];
#pragma unroll
                for (int i = 0; i < DIAG_PANEL; ++i) {
                    double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < DIAG_PANEL; ++k)
                        if (k < i) acc = fma(-D_(i, k), (k >= c) ? x[k] : 0.0, acc);
                    x[i] = (i >= c) ? acc * dinv[i] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < DIAG_PANEL; ++i) {
                    I_(i, c) = x[i];
                    if (i >= c) W_(j0 + i, j0 + c) = x[i];
                }
            }
        }
        __syncthreads();
        const int t0 = j0 + DIAG_PANEL;
        if (t0 < POTRF_NB) {
            // ---- (b) panel rows below: P <- P * inv(D)'   (P(i, c) = sum_k Pold(i, k) I(c, k), k <= c) ----
            // each thread owns whole rows so that the in-place product needs no second buffer
            for (int i = t0 + tid; i < POTRF_NB; i += DIAG_THREADS) {
                double pr[DIAG_PANEL];
#pragma unroll
                for (int k = 0; k < DIAG_PANEL; ++k) pr[k] = T_(i, j0 + k);
#pragma unroll
                for (int c = 0; c < DIAG_PANEL; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < DIAG_PANEL; ++k)
                        if (k <= c) acc = fma(pr[k], I_(c, k), acc);
                    T_(i, j0 + c) = acc;
                }
            }
            __syncthreads();
            // ---- (c) trailing update T(t0.., t0..) -= P P' (lower tiles) ----
            smem_tile_gemm_nt<true, true>(T, t0, t0, POTRF_NB - t0, POTRF_NB - t0, T, t0, j0, T, t0, j0, DIAG_PANEL, -1.0, tid);
            __syncthreads();
        }
    }
    {
        const int i = tid & (POTRF_NB - 1);
        if (i < nb)
            for (int k = tid >> 7; k < nb; k += DIAG_THREADS / POTRF_NB)
                A[i + (long)k * ld] = T_(i, k);        // strict upper of the block becomes 0
    }
    // ---- blocked triangular inverse: W holds inv of the 16 x 16 diagonal blocks; merge pairs 16 -> 32 -> 64 -> 128.
    //      X = L21 * W11 goes through the (now free) strict upper part of T: block (o.., o+b..)  ----
    for (int b = DIAG_PANEL; b < POTRF_NB; b *= 2) {
        // X(o+b.., o..) = L21 * W11 for every pair, stored transposed-free in T's upper block region (rows o.., cols o+b..)
        for (int o = 0; o + b < POTRF_NB; o += 2 * b) {
            // Xt(c, r) := X(r, c) cannot alias the lower part; keep X in T(o + c, o + b + r) = upper block (o.., o+b..)
            // computed as Xt = W11' * L21'  ==  C(i = c, j = r) = sum_k W11(k, c) L21(r, k): NT form needs A(i,k) = W11(k,i).
            // Simpler: compute X directly into the upper block with a plain loop tiling over (r, c).
            const int tr = b >> 2, tc = b >> 2;
            for (int tile = tid; tile < tr * tc; tile += DIAG_THREADS) {
                const int r0 = (tile % tr) * 4, c0 = (tile / tr) * 4;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) acc[a][bb] = 0.0;
                for (int k = c0; k < b; ++k) {                       // W11(k, c) = 0 for k < c
                    double lv[4], wv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { lv[a] = T_(o + b + r0 + a, o + k); wv[a] = WG_(o + k, o + c0 + a); }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) acc[a][bb] = fma(lv[a], wv[bb], acc[a][bb]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) T_(o + c0 + bb, o + b + r0 + a) = acc[a][bb];   // X(r, c) kept at T(c, b + r)
            }
        }
        __syncthreads();
        for (int o = 0; o + b < POTRF_NB; o += 2 * b) {
            // W21(r, c) = - sum_k W22(r, k) X(k, c),  W22(r, k) = 0 for k > r
            const int tr = b >> 2, tc = b >> 2;
            for (int tile = tid; tile < tr * tc; tile += DIAG_THREADS) {
                const int r0 = (tile % tr) * 4, c0 = (tile / tr) * 4;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) acc[a][bb] = 0.0;
                for (int k = 0; k < r0 + 4; ++k) {
                    double wv[4], xv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { wv[a] = WG_(o + b + r0 + a, o + b + k); xv[a] = T_(o + c0 + a, o + b + k); }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) acc[a][bb] = fma(wv[a], xv[bb], acc[a][bb]);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) W_(o + b + r0 + a, o + c0 + bb) = -acc[a][bb];
            }
        }
        __syncthreads();
    }
    {
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB)
            inv[i + (long)k * POTRF_NB] = (i < nb && k < nb && i >= k) ? W_(i, k) : 0.0;
    }
#undef WG_
    {
    }
#undef T_
#undef W_
#undef D_
#undef I_
}

// 1/sqrt(d) for d > 0 without the library's FP64 rsqrt/div sequences on the serial pivot chain:
// single-precision MUFU.RSQ seed (22 bits) + two Newton steps in double (-> < 1 ulp).
__device__ __forceinline__ double rsqrt_pivot(double d) {
    if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
    double r = (double)rsqrtf((float)d);
    const double h = 0.5 * d;
    double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    return r;
}

// v2 of the diagonal-block kernel.  Same contract as v1 (factor nb x nb <= 128 in shared memory, write L back, write
// inv(L) to `inv`), restructured so that the serial chain is a few hundred cycles per column and every O(n³)
// piece runs on the FP64 tensor pipe (DMMA m8n8k4 from shared-memory fragments, pitch 132 = conflict-free):
//  * 16-column panels; inside a panel thread i (< 128) owns ROW i of the panel in registers and ALL rows below the
//    pivot take part in the column-by-column elimination (the panel's TRSM is folded into its factorisation -- no
//    per-panel 16 x 16 inverse on the chain).  Per column the 16 diagonal-block rows publish their unscaled entries
//    u_k = T(j0+k, j0+c) (one 128-thread named barrier), every row derives 1/sqrt(u_c) itself and applies
//    row[k] -= (row[c] r)(u_k r).
//  * rank-16 trailing update: 16 x 16 output tiles dealt to the 16 warps, 2 x 2 DMMA tiles each.
//  * inverse afterwards: the eight 16 x 16 diagonal inverses in parallel (one warp each), then recursive doubling
//    16 -> 32 -> 64 -> 128 with X = L21 W11 and W21 = -W22 X as warp-level DMMA products.  W = inv(L) is kept
//    TRANSPOSED in T's strict upper triangle (W(i,k), i > k, at T(k,i)), its diagonal in dinv[].
constexpr int XS_PITCH = 68;   // X scratch: column c of pair q at Xs[(q*b + c) * XS_PITCH + r]
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* A, long ld, int nb, double* inv, int* fail, int base, long long* dbg) {
    extern __shared__ double dsm[];
    double* T = dsm;                                   // T(i,k) at k*DIAG_PITCH + i
    double* Xs = dsm + POTRF_NB * DIAG_PITCH;          // 64 x XS_PITCH
    double* ubuf = Xs + 64 * XS_PITCH;                 // 2 x 16 published column entries
    double* dinv = ubuf + 2 * DIAG_PANEL;              // 128 reciprocal diagonal entries of L = diagonal of inv(L)
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark = clock64();
#define DIAG_PHASE(i) do { if (dbg) { const long long t_ = clock64(); tk[i] += t_ - tmark; tmark = t_; } } while (0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;           // DMMA fragment coordinates
#define T_(i, k) T[(k) * DIAG_PITCH + (i)]
// W(i,k) = inv(L)(i,k) for i >= k
#define WL_(i, k) (((i) > (k)) ? T_((k), (i)) : (((i) == (k)) ? dinv[(i)] : 0.0))
    {
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB)
            T_(i, k) = (i < nb && k < nb) ? ((i >= k) ? A[i + (long)k * ld] : 0.0) : (i == k ? 1.0 : 0.0);
    }
    __syncthreads();
    DIAG_PHASE(0);
    for (int jb = 0; jb < DIAG_NPANEL; ++jb) {
        const int j0 = jb * DIAG_PANEL;
        if (tid < POTRF_NB) {
            // ---- panel factorisation + TRSM: thread i owns row i (rows < j0 only keep the barrier count) ----
            const int i = tid;
            const int rl = i - j0;                       // row inside the panel's diagonal block if 0 <= rl < 16
            const bool active = i >= j0;
            double row[DIAG_PANEL];
#pragma unroll
            for (int c = 0; c < DIAG_PANEL; ++c) row[c] = (active && (rl >= DIAG_PANEL || c <= rl)) ? T_(i, j0 + c) : 0.0;
#pragma unroll
            for (int c = 0; c < DIAG_PANEL; ++c) {
                double* ub = ubuf + (c & 1) * DIAG_PANEL;
                if (rl >= c && rl < DIAG_PANEL) ub[rl] = row[c];          // u_k, k = c..15 (row k of the diagonal block)
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                double d = ub[c];
                if (!(d > 0.0)) {                                         // also catches NaN
                    if (i == j0 + c && j0 + c < nb) atomicCAS(fail, 0, base + j0 + c + 1);
                    d = 1.0;
                }
                const double r = rsqrt_pivot(d);
                if (i == j0 + c) dinv[j0 + c] = r;
                const double lcv = (rl == c) ? d * r : row[c] * r;        // L(i, j0+c)
                row[c] = lcv;
#pragma unroll
                for (int k = c + 1; k < DIAG_PANEL; ++k) row[k] = fma(-lcv, ub[k] * r, row[k]);
            }
            if (active) {
#pragma unroll
                for (int c = 0; c < DIAG_PANEL; ++c)
                    if (rl >= DIAG_PANEL || c <= rl) T_(i, j0 + c) = row[c];
            }
        }
        __syncthreads();
        DIAG_PHASE(1);
        const int t0 = j0 + DIAG_PANEL;
        if (t0 < POTRF_NB) {
            // ---- trailing update T(t0.., t0..) -= P P' on the lower 16 x 16 tiles (DMMA, K = 16) ----
            const int nb16 = (POTRF_NB - t0) / 16;
            const int ntile = nb16 * (nb16 + 1) / 2;
            for (int t = warp; t < ntile; t += DIAG_THREADS / 32) {
                int bi = 0, rem = t;
                while (rem > bi) { rem -= bi + 1; ++bi; }
                const int m0 = t0 + 16 * bi, n0 = t0 + 16 * rem;
                double acc[2][2][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                        for (int e = 0; e < 2; ++e) acc[mi][nj][e] = T_(m0 + mi * 8 + lr, n0 + nj * 8 + 2 * lc + e);
#pragma unroll
                for (int k4 = 0; k4 < DIAG_PANEL; k4 += 4) {
                    double af[2], bf[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        af[q] = -T_(m0 + q * 8 + lr, j0 + k4 + lc);
                        bf[q] = T_(n0 + q * 8 + lr, j0 + k4 + lc);
                    }
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int nj = 0; nj < 2; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int m = m0 + mi * 8 + lr, n = n0 + nj * 8 + 2 * lc + e;
                            if (m >= n) T_(m, n) = acc[mi][nj][e];
                        }
            }
            __syncthreads();
            DIAG_PHASE(2);
        }
    }
    {
        const int i = tid & (POTRF_NB - 1);
        if (i < nb)
            for (int k = tid >> 7; k < nb; k += DIAG_THREADS / POTRF_NB)
                A[i + (long)k * ld] = (i >= k) ? T_(i, k) : 0.0;
    }
    DIAG_PHASE(3);
    // ---- 16 x 16 diagonal inverses: warp w < 8 takes block w, lane c < 16 computes column c by forward substitution ----
    if (warp < DIAG_NPANEL && lane < DIAG_PANEL) {
        const int j0 = warp * DIAG_PANEL, c = lane;
        double x[DIAG_PANEL];
#pragma unroll
        for (int i = 0; i < DIAG_PANEL; ++i) {
            double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < DIAG_PANEL; ++k)
                if (k < i) acc = fma(-T_(j0 + i, j0 + k), (k >= c) ? x[k] : 0.0, acc);
            x[i] = (i >= c) ? acc * dinv[j0 + i] : 0.0;
        }
        __syncwarp(0x0000ffffu);     // every lane has finished reading L before the block's upper triangle is overwritten
#pragma unroll
        for (int i = 0; i < DIAG_PANEL; ++i)
            if (i > c) T_(j0 + c, j0 + i) = x[i];          // W(j0+i, j0+c), transposed; the diagonal x[c] equals dinv
    }
    __syncthreads();
    DIAG_PHASE(4);
    // ---- blocked triangular inverse: merge pairs 16 -> 32 -> 64 -> 128 ----
    for (int b = DIAG_PANEL; b < POTRF_NB; b *= 2) {
        const int nt = b / 16;                       // 16 x 16 tiles per side of a b x b block
        const int njobs = (POTRF_NB / (2 * b)) * nt * nt;   // 4, 8, 16
        // X = L21 W11  (W11 lower: k >= c)
        if (warp < njobs) {
            const int pair = warp / (nt * nt), ri = (warp / nt) % nt, ci = warp % nt;
            const int o = pair * 2 * b;
            double acc[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
            for (int k4 = 16 * ci; k4 < b; k4 += 4) {
                double af[2], bf[2];
                const int k = k4 + lc;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    af[q] = T_(o + b + 16 * ri + q * 8 + lr, o + k);
                    const int c = 16 * ci + q * 8 + lr;
                    bf[q] = WL_(o + k, o + c);
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int nj = 0; nj < 2; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = 16 * ri + mi * 8 + lr, c = 16 * ci + nj * 8 + 2 * lc + e;
                        Xs[(pair * b + c) * XS_PITCH + r] = acc[mi][nj][e];
                    }
        }
        __syncthreads();
        // W21 = -W22 X  (W22 lower: k <= r), stored transposed into T's upper triangle
        if (warp < njobs) {
            const int pair = warp / (nt * nt), ri = (warp / nt) % nt, ci = warp % nt;
            const int o = pair * 2 * b;
            double acc[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
            for (int k4 = 0; k4 < 16 * ri + 16; k4 += 4) {
                double af[2], bf[2];
                const int k = k4 + lc;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int r = 16 * ri + q * 8 + lr;
                    af[q] = WL_(o + b + r, o + b + k);
                    bf[q] = Xs[(pair * b + 16 * ci + q * 8 + lr) * XS_PITCH + k];
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int nj = 0; nj < 2; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = 16 * ri + mi * 8 + lr, c = 16 * ci + nj * 8 + 2 * lc + e;
                        T_(o + c, o + b + r) = -acc[mi][nj][e];        // W(o+b+r, o+c)
                    }
        }
        __syncthreads();
    }
    DIAG_PHASE(5);
    {
        // inv(i,k) = W(i,k): coalesced global stores (i fastest); the transposed shared-memory reads are 4-way conflicted, once
        const int i = tid & (POTRF_NB - 1);
        for (int k = tid >> 7; k < POTRF_NB; k += DIAG_THREADS / POTRF_NB)
            inv[i + (long)k * POTRF_NB] = (i < nb && k < nb && i >= k) ? WL_(i, k) : 0.0;
    }
    DIAG_PHASE(6);
    if (dbg && tid == 0)
        for (int i = 0; i < 7; ++i) atomicAdd((unsigned long long*)dbg + i, (unsigned long long)tk[i]);
#undef DIAG_PHASE
#undef WL_
#undef T_
}

int potrf_lower(kb_ctx* ctx, double* A, long ld, int p, double* invdiag, int* d_fail) {
    static const bool use_v1 = [] { const char* e = getenv("KB_POTRF_DIAG"); return e && e[0] == '1'; }();
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)potrf_diag_kernel, DIAG_SMEM));
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)potrf_diag_kernel_v1, DIAG_SMEM));
    // KB_POTRF_PHASES=1: per-phase cycle counts of the diagonal kernel (thread 0's clock64), printed per factorisation
    static const bool dbg_on = getenv("KB_POTRF_PHASES") != nullptr;
    long long* dbg = nullptr;
    if (dbg_on) {
        cudaMalloc(&dbg, 8 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 8 * sizeof(long long), ctx->stream);
    }
    std::vector<cudaEvent_t> ev;
    auto mark = [&]() {
        if (!dbg_on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        ev.push_back(e);
    };
    // Right-looking blocked Cholesky with ONE panel of look-ahead.  The trailing update of panel b is split into the
    // next block column (on the main stream) and the rest (on ctx->stream_aux); the single-CTA diagonal kernel and the
    // panel solve of block b + 1 then run while the rest of update b is still streaming, so the serial diagonal chain
    // (2.3 of 5.3 ms at p = 5000) no longer leaves the other 147 SMs idle.  KB_POTRF_LOOKAHEAD=0 restores the plain order.
    static const bool lookahead = [] { const char* e = getenv("KB_POTRF_LOOKAHEAD"); return !(e && e[0] == '0'); }() && !dbg_on;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream_aux;
    bool aux_pending = false;
    int blk = 0;
    for (int r0 = 0; r0 < p; r0 += POTRF_NB, ++blk) {
        const int nb = (p - r0 < POTRF_NB) ? p - r0 : POTRF_NB;
        double* Abb = A + r0 + (long)r0 * ld;
        double* inv = invdiag + (size_t)blk * POTRF_NB * POTRF_NB;
        mark();
        if (use_v1) potrf_diag_kernel_v1<<<1, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(Abb, ld, nb, inv, d_fail, r0);
        else potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(Abb, ld, nb, inv, d_fail, r0, dbg);
        ctx->launches++;
        KB_CUDA(ctx, cudaGetLastError());
        mark();
        const int rem = p - r0 - nb;
        if (rem <= 0) break;
        double* A21 = A + (r0 + nb) + (long)r0 * ld;
        // L21 = A21 * inv(L11)'  IN PLACE.  inv(L11)' is upper triangular, so output columns [n0, n0 + BN) depend on input
        // columns k < n0 + BN only: the N tiles are launched one by one from the LAST to the first, each reading only
        // columns that no earlier launch has overwritten (a CTA reads and writes its own rows only).  A single launch over
        // several N tiles would let the tile of columns [0, BN) overwrite what the next tile still has to read.
        for (int n0 = ((nb - 1) / GemmDefaultCfg::BN) * GemmDefaultCfg::BN; n0 >= 0; n0 -= GemmDefaultCfg::BN) {
            const int nw = (nb - n0 < GemmDefaultCfg::BN) ? nb - n0 : GemmDefaultCfg::BN;
            KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, nw, n0 + nw, 1.0, A21, ld, inv + n0, POTRF_NB, 0.0,
                         A21 + (long)n0 * ld, ld));
        }
        mark();
        // A22 -= L21 L21'  (lower tiles)
        double* A22 = A + (r0 + nb) + (long)(r0 + nb) * ld;
        const int nb2 = rem < POTRF_NB ? rem : POTRF_NB;          // width of the next block column
        if (!lookahead || rem <= nb2) {
            if (aux_pending) { KB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->aux_ev[1], 0)); aux_pending = false; }
            KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, rem, nb, -1.0, A21, ld, A21, ld, 1.0, A22, ld, 0, 1, 0));
            continue;
        }
        // (a) next block column on the main stream -- after the previous panel's aux update, which touched it too
        if (aux_pending) KB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->aux_ev[1], 0));
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, rem, nb2, nb, -1.0, A21, ld, A21, ld, 1.0, A22, ld, 0, 1, 0));
        KB_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], s0));
        // (b) the remaining columns on the aux stream (needs L21 of this panel: ordered by aux_ev[0])
        KB_CUDA(ctx, cudaStreamWaitEvent(s1, ctx->aux_ev[0], 0));
        ctx->stream = s1;
        const int rc = gemm1(ctx, A_MCONTIG, B_NCONTIG, rem - nb2, rem - nb2, nb, -1.0, A21 + nb2, ld, A21 + nb2, ld, 1.0,
                             A22 + nb2 + (long)nb2 * ld, ld, 0, 1, 0);
        ctx->stream = s0;
        KB_TRY(rc);
        KB_CUDA(ctx, cudaEventRecord(ctx->aux_ev[1], s1));
        aux_pending = true;
    }
    if (aux_pending) KB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->aux_ev[1], 0));
    if (dbg) {
        long long h[8];
        cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(dbg);
        double t[3] = {0, 0, 0};
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            t[i % 3] += ms;
        }
        fprintf(stderr, "[potrf p=%d, ms] diag=%.3f trsm=%.3f syrk=%.3f\n", p, t[0], t[1], t[2]);
        for (auto& e : ev) cudaEventDestroy(e);
        fprintf(stderr, "[potrf_diag x%d, kcycles] load=%lld panel=%lld trailing=%lld storeL=%lld inv16=%lld doubling=%lld storeInv=%lld\n",
                blk + 1, h[0] / 1000, h[1] / 1000, h[2] / 1000, h[3] / 1000, h[4] / 1000, h[5] / 1000, h[6] / 1000);
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
