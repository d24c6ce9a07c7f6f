// Hand-written FP64 tensor-core GEMM for sm_100a.
//
// C (M x N, column-major) = alpha * sum_seg A_seg * B_seg + beta * Cin + bias_n
//
// The FP64 tensor path on Blackwell is the warp-level DMMA: `mma.sync.aligned.m8n8k4.f64`
// (SASS DMMA.8x8x4; tcgen05 has no f64 kind, and the m16n8k* f64 shapes are split into 8x8x4 by
// ptxas on sm_100a), fed from shared memory.  Tiles are staged global -> shared with a 3-stage
// cp.async (LDGSTS) pipeline, 16-byte transactions, padded rows so that every fragment load is
// bank-conflict free.
//
// Default CTA tile 64 x 64 x 16, 128 threads = 4 warps laid out 2 (M) x 2 (N), a warp owns 32 x 32 = 4 x 4 DMMA tiles
// (32 accumulator doubles per thread), 3 CTAs per SM (GemmCfg below; the 128 x 128 / 8-warp tile of early round 1 and
// the 64 x 128 tile are kept as build-time variants, profiles/r01_gemm_variants.md).  A TMA-staged variant of the same
// inner loop (cp.async.bulk.tensor + mbarrier ring, 128-byte swizzle) lives in dgemm_tma.cu; profiles/r02_gemm_variants.md
// has the A/B measurement.
//
// Up to 3 K-segments (A_i, B_i) are accumulated into one tile so that the sampling stage
// X̃ = X − X·ΣinvS + Z·L (+ T·M) is ONE fused GEMM.  Each segment may declare triangular
// structure (k-range clipped per tile) so triangular operands cost half the flops.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace kb {

// CTA tile configurations (BN = 128, BK = 16, 3-stage cp.async pipeline in all of them):
//   GemmCfg<128, 128, 2, 4, 1>  128 x 128, 8 warps of 64 x 32, 1 CTA / SM      (first kernel of round 1)
//   GemmCfg< 64, 128, 2, 2, 2>   64 x 128, 4 warps of 32 x 64, 2 CTAs / SM: the two resident CTAs run out of phase, so
//                                one feeds the DMMA pipe while the other sits at its k-tile barrier / waits for cp.async
//   GemmCfg< 64,  64, 2, 2, 3>   64 x  64, 4 warps of 32 x 32, 3 CTAs / SM
#ifndef KB_GEMM_BK
#define KB_GEMM_BK 16
#endif
#ifndef KB_GEMM_STAGES
#define KB_GEMM_STAGES 3
#endif
constexpr int GEMM_BK = KB_GEMM_BK;
constexpr int GEMM_STAGES = KB_GEMM_STAGES;
constexpr int GEMM_PADK = GEMM_BK + 4;   // [m][k] rows: stride mod 16 == 4 -> conflict-free fragments

template <int BM_, int BN_, int WM_, int WN_, int MIN_CTAS_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_;
    static constexpr int WARPS = WM * WN, THREADS = 32 * WARPS;
    static constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;          // 8 x 8 DMMA tiles per warp
    static constexpr int PADM = BM + 4, PADN = BN + 4;                // [k][m] / [k][n] rows: stride mod 16 == 4
    static constexpr int A_TILE = (GEMM_BK * PADM > BM * GEMM_PADK) ? GEMM_BK * PADM : BM * GEMM_PADK;
    static constexpr int B_TILE = (GEMM_BK * PADN > BN * GEMM_PADK) ? GEMM_BK * PADN : BN * GEMM_PADK;
    static constexpr int STAGE = A_TILE + B_TILE;
    static constexpr size_t SMEM_BYTES = (size_t)GEMM_STAGES * STAGE * sizeof(double);
    static constexpr int MIN_CTAS = MIN_CTAS_;
};

enum : int { A_MCONTIG = 0, A_KCONTIG = 1 };   // A(m,k) at m + k*lda   |   k + m*lda
enum : int { B_KCONTIG = 0, B_NCONTIG = 1 };   // B(k,n) at k + n*ldb   |   n + k*ldb

// per-segment k-range clipping (operand known to be zero outside)
enum : int {
    SEG_KLO_N = 1,   // k >= n0            (B lower-triangular in (k,n):  B[k,n] = 0 for k < n)
    SEG_KLO_M = 2,   // k >= m0            (A upper-triangular in (m,k):  A[m,k] = 0 for k < m)
    SEG_KHI_N = 4,   // k <  n0 + BN       (B upper-triangular in (k,n))
    SEG_KHI_M = 8,   // k <  m0 + BM       (A lower-triangular in (m,k))
};

struct GemmSeg {
    const double* A;
    long lda;
    const double* B;
    long ldb;
    int K;
    int flags;
};

struct GemmParams {
    int M, N;
    int nseg;
    GemmSeg seg[3];
    double* C;
    long ldc;
    const double* Cin;   // may alias C (each element is read then written by the same thread)
    long ldcin;
    double alpha, beta;
    const double* bias_n;   // optional, added per column n
    int tile_mode;          // 0 all tiles, 1 lower tiles only (skip tiles with m0 + BM <= n0)
    int store_mode;         // 0 all, 1 only elements with m >= n (lower incl. diagonal)
    double* C2;             // optional second destination: columns n >= nsplit are stored to C2 + (n - nsplit) * ldc
    int nsplit;             //   (C2 == nullptr: off).  Requires beta == 0 or Cin laid out like the logical C.
    // Batched launch (gridDim.z = batch > 1): problem z uses seg[s].A + z * zA[s], seg[s].B + z * zB[s], C + z * zC
    // (Cin / bias are shared); the LAST problem uses only its first `last_nseg` segments when last_nseg > 0.
    // One launch for the m copies of the sampling stage: 5 x 5056 CTAs instead of 5 launches with a partial last
    // wave each.
    int batch;
    int last_nseg;
    int khi_slack;          // SEG_KHI_N segments: B is zero for k >= n + 1 + khi_slack (banded below the diagonal)
    long zA[3], zB[3];
    long zC;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stage one BM x BK tile of A and one BK x BN tile of B into shared memory.
template <class Cfg, int AL, int BL, bool VEC>
__device__ __forceinline__ void gemm_load_stage(const GemmSeg& sg, int M, int N, int m0, int n0, int k0,
                                                int kend, double* As, double* Bs, int tid) {
    // ---- A tile ----
    if (AL == A_MCONTIG) {
        if (VEC) {
#pragma unroll
            for (int it = 0; it < (Cfg::BM * GEMM_BK / 2) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int k = c / (Cfg::BM / 2), m = (c % (Cfg::BM / 2)) * 2;
                int gm = m0 + m, gk = k0 + k;
                int valid = (gk < kend) ? max(0, min(2, M - gm)) : 0;
                const double* src = valid ? sg.A + (long)gm + (long)gk * sg.lda : sg.A;
                cp_async16(As + k * Cfg::PADM + m, src, valid * 8);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (Cfg::BM * GEMM_BK) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int k = c / Cfg::BM, m = c % Cfg::BM;
                int gm = m0 + m, gk = k0 + k;
                int valid = (gk < kend && gm < M);
                const double* src = valid ? sg.A + (long)gm + (long)gk * sg.lda : sg.A;
                cp_async8(As + k * Cfg::PADM + m, src, valid * 8);
            }
        }
    } else {
        if (VEC) {
#pragma unroll
            for (int it = 0; it < (Cfg::BM * GEMM_BK / 2) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int m = c / (GEMM_BK / 2), k = (c % (GEMM_BK / 2)) * 2;
                int gm = m0 + m, gk = k0 + k;
                int valid = (gm < M) ? max(0, min(2, kend - gk)) : 0;
                const double* src = valid ? sg.A + (long)gk + (long)gm * sg.lda : sg.A;
                cp_async16(As + m * GEMM_PADK + k, src, valid * 8);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (Cfg::BM * GEMM_BK) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int m = c / GEMM_BK, k = c % GEMM_BK;
                int gm = m0 + m, gk = k0 + k;
                int valid = (gk < kend && gm < M);
                const double* src = valid ? sg.A + (long)gk + (long)gm * sg.lda : sg.A;
                cp_async8(As + m * GEMM_PADK + k, src, valid * 8);
            }
        }
    }
    // ---- B tile ----
    if (BL == B_KCONTIG) {
        if (VEC) {
#pragma unroll
            for (int it = 0; it < (Cfg::BN * GEMM_BK / 2) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int n = c / (GEMM_BK / 2), k = (c % (GEMM_BK / 2)) * 2;
                int gn = n0 + n, gk = k0 + k;
                int valid = (gn < N) ? max(0, min(2, kend - gk)) : 0;
                const double* src = valid ? sg.B + (long)gk + (long)gn * sg.ldb : sg.B;
                cp_async16(Bs + n * GEMM_PADK + k, src, valid * 8);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (Cfg::BN * GEMM_BK) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int n = c / GEMM_BK, k = c % GEMM_BK;
                int gn = n0 + n, gk = k0 + k;
                int valid = (gk < kend && gn < N);
                const double* src = valid ? sg.B + (long)gk + (long)gn * sg.ldb : sg.B;
                cp_async8(Bs + n * GEMM_PADK + k, src, valid * 8);
            }
        }
    } else {
        if (VEC) {
#pragma unroll
            for (int it = 0; it < (Cfg::BN * GEMM_BK / 2) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int k = c / (Cfg::BN / 2), n = (c % (Cfg::BN / 2)) * 2;
                int gn = n0 + n, gk = k0 + k;
                int valid = (gk < kend) ? max(0, min(2, N - gn)) : 0;
                const double* src = valid ? sg.B + (long)gn + (long)gk * sg.ldb : sg.B;
                cp_async16(Bs + k * Cfg::PADN + n, src, valid * 8);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (Cfg::BN * GEMM_BK) / Cfg::THREADS; ++it) {
                int c = tid + it * Cfg::THREADS;
                int k = c / Cfg::BN, n = c % Cfg::BN;
                int gn = n0 + n, gk = k0 + k;
                int valid = (gk < kend && gn < N);
                const double* src = valid ? sg.B + (long)gn + (long)gk * sg.ldb : sg.B;
                cp_async8(Bs + k * Cfg::PADN + n, src, valid * 8);
            }
        }
    }
}


// Interior-tile fast path of the stage loader (VEC layouts, tile fully inside M x N, k-tile fully inside the segment's
// k-range): every 16-byte chunk of a thread sits a fixed stride behind the previous one, so the whole stage is 8
// unpredicated LDGSTS from two running pointers -- the general loader above spends ~180 integer instructions per
// k-tile on index / bounds arithmetic (SASS of r01e), which a microbenchmark of the same LDS + DMMA + barrier loop
// without global loads (tests/ub_dmma_smem.cu: 35.3 TFLOP/s) showed to be where the kernel loses against cuBLAS.
template <class Cfg, int AL, int BL>
struct FastLoad {
    const double* a;      // this thread's first A chunk of the current k-tile
    const double* b;
    long a_it, b_it;      // element stride between a thread's consecutive chunks
    long a_kt, b_kt;      // element stride between consecutive k-tiles
    int sa, sb;           // shared-memory offsets (doubles) of the first chunks
    static constexpr int ITA = (Cfg::BM * GEMM_BK / 2) / Cfg::THREADS, ITB = (Cfg::BN * GEMM_BK / 2) / Cfg::THREADS;
    static constexpr int SA_IT = (AL == A_MCONTIG) ? (Cfg::THREADS / (Cfg::BM / 2)) * Cfg::PADM : (Cfg::THREADS / (GEMM_BK / 2)) * GEMM_PADK;
    static constexpr int SB_IT = (BL == B_KCONTIG) ? (Cfg::THREADS / (GEMM_BK / 2)) * GEMM_PADK : (Cfg::THREADS / (Cfg::BN / 2)) * Cfg::PADN;
    static constexpr bool OK = (Cfg::THREADS % (Cfg::BM / 2) == 0) && (Cfg::THREADS % (Cfg::BN / 2) == 0) && (Cfg::THREADS % (GEMM_BK / 2) == 0);
    __device__ __forceinline__ void init(const GemmSeg& sg, int m0, int n0, int k0, int tid) {
        if (AL == A_MCONTIG) {
            const int k = tid / (Cfg::BM / 2), m = (tid % (Cfg::BM / 2)) * 2;
            a = sg.A + (long)(m0 + m) + (long)(k0 + k) * sg.lda;
            a_it = (long)(Cfg::THREADS / (Cfg::BM / 2)) * sg.lda; a_kt = (long)GEMM_BK * sg.lda;
            sa = k * Cfg::PADM + m;
        } else {
            const int m = tid / (GEMM_BK / 2), k = (tid % (GEMM_BK / 2)) * 2;
            a = sg.A + (long)(k0 + k) + (long)(m0 + m) * sg.lda;
            a_it = (long)(Cfg::THREADS / (GEMM_BK / 2)) * sg.lda; a_kt = GEMM_BK;
            sa = m * GEMM_PADK + k;
        }
        if (BL == B_KCONTIG) {
            const int n = tid / (GEMM_BK / 2), k = (tid % (GEMM_BK / 2)) * 2;
            b = sg.B + (long)(k0 + k) + (long)(n0 + n) * sg.ldb;
            b_it = (long)(Cfg::THREADS / (GEMM_BK / 2)) * sg.ldb; b_kt = GEMM_BK;
            sb = n * GEMM_PADK + k;
        } else {
            const int k = tid / (Cfg::BN / 2), n = (tid % (Cfg::BN / 2)) * 2;
            b = sg.B + (long)(n0 + n) + (long)(k0 + k) * sg.ldb;
            b_it = (long)(Cfg::THREADS / (Cfg::BN / 2)) * sg.ldb; b_kt = (long)GEMM_BK * sg.ldb;
            sb = k * Cfg::PADN + n;
        }
    }
    // loads k-tile number `kt` (relative to the k0 given to init) into the stage buffers
    __device__ __forceinline__ void load(int kt, double* As, double* Bs) const {
        const double* pa = a + (long)kt * a_kt;
        const double* pb = b + (long)kt * b_kt;
#pragma unroll
        for (int it = 0; it < ITA; ++it) cp_async16(As + sa + it * SA_IT, pa + it * a_it, 16);
#pragma unroll
        for (int it = 0; it < ITB; ++it) cp_async16(Bs + sb + it * SB_IT, pb + it * b_it, 16);
    }
};

template <class Cfg, int AL, int BL, bool VEC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) dgemm_dmma_kernel(const GemmParams P) {
    extern __shared__ __align__(16) double gemm_smem[];
    constexpr int BM = Cfg::BM, BN = Cfg::BN, MI = Cfg::MI, NJ = Cfg::NJ;
    constexpr int WTM = MI * 8, WTN = NJ * 8;          // warp tile
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
    // Rasterisation.  Full grids: GROUP_M consecutive M tiles share one pass over the N tiles, so a wave of resident
    // CTAs covers a ~square patch of C and the k-slices of A and B it streams are shared through L2 by ~20 CTAs each
    // (x-fastest order made a wave 64 M tiles x 7 N tiles: every wave re-read all of A from DRAM, ncu r01b: 5.8 GB of
    // traffic per sampling GEMM vs 1.1 GB algorithmic).  Lower-tile grids keep the plain order.
    int bx = blockIdx.x, by = blockIdx.y;
    if (P.tile_mode == 0) {
        constexpr int GROUP_M = 16;
        const int num_m = gridDim.x, num_n = gridDim.y;
        const int pid = blockIdx.x + blockIdx.y * num_m;
        const int per_group = GROUP_M * num_n;
        const int first_m = (pid / per_group) * GROUP_M;
        const int gsz = min(num_m - first_m, GROUP_M);
        const int in_group = pid % per_group;
        bx = first_m + in_group % gsz;
        by = in_group / gsz;
    }
    const int m0 = bx * BM, n0 = by * BN;
    if (P.tile_mode == 1 && m0 + BM <= n0) return;
    const long zb = blockIdx.z;
    const int nseg = (P.batch > 1 && P.last_nseg > 0 && (int)zb == P.batch - 1) ? P.last_nseg : P.nseg;
    double* const Cz = P.C + zb * P.zC;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int lr = lane >> 2, lc = lane & 3;

    // beta * Cin enters through the accumulators: all of a thread's Cin loads are issued here, back to back and
    // ahead of the first k-tile (they overlap the cp.async prologue).  Loading in the epilogue instead serialises
    // one DRAM round trip per element, because Cin may alias C and every load would have to wait for the
    // previous store -- that made short-K updates (rank-64/128 passes, Cholesky trailing updates) latency-bound.
    const bool preload = (P.beta != 0.0 && P.alpha != 0.0);
    if (preload) {
        const double r = P.beta / P.alpha;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + wm * WTM + i * 8 + lr;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + wn * WTN + j * 8 + 2 * lc + e;
                    if (m < P.M && n < P.N && !(P.store_mode == 1 && m < n))
                        acc[i][j][e] = r * P.Cin[(long)m + (long)n * P.ldcin];
                }
            }
        }
    }

    for (int s = 0; s < nseg; ++s) {
        GemmSeg sg = P.seg[s];
        sg.A += zb * P.zA[s];
        sg.B += zb * P.zB[s];
        int kbeg = 0, kend = sg.K;
        if (sg.flags & SEG_KLO_N) kbeg = max(kbeg, n0);
        if (sg.flags & SEG_KLO_M) kbeg = max(kbeg, m0);
        if (sg.flags & SEG_KHI_N) kend = min(kend, n0 + BN + P.khi_slack);
        if (sg.flags & SEG_KHI_M) kend = min(kend, m0 + BM);
        kbeg = (kbeg / GEMM_BK) * GEMM_BK;
        if (kend <= kbeg) continue;
        const int nkt = (kend - kbeg + GEMM_BK - 1) / GEMM_BK;

        __syncthreads();   // previous segment's readers are done with the stage buffers
        // interior tiles take the unpredicated fast loader for every k-tile that lies fully inside [kbeg, kend)
        FastLoad<Cfg, AL, BL> fl;
        const bool fast = VEC && FastLoad<Cfg, AL, BL>::OK && (m0 + BM <= P.M) && (n0 + BN <= P.N);
        const int nfull = fast ? (kend - kbeg) / GEMM_BK : 0;      // k-tiles [0, nfull) are complete
        if (fast) fl.init(sg, m0, n0, kbeg, tid);
        auto load_tile = [&](int t) {
            double* As_ = gemm_smem + (size_t)(t % GEMM_STAGES) * Cfg::STAGE;
            if (t < nfull) fl.load(t, As_, As_ + Cfg::A_TILE);
            else gemm_load_stage<Cfg, AL, BL, VEC>(sg, P.M, P.N, m0, n0, kbeg + t * GEMM_BK, kend, As_, As_ + Cfg::A_TILE, tid);
        };
        // prologue
#pragma unroll
        for (int st = 0; st < GEMM_STAGES - 1; ++st) {
            if (st < nkt) load_tile(st);
            cp_async_commit();
        }
        for (int kt = 0; kt < nkt; ++kt) {
            cp_async_wait<GEMM_STAGES - 2>();
            __syncthreads();
            // prefetch tile kt + STAGES - 1 into the buffer consumed at iteration kt - 1
            {
                int nk = kt + GEMM_STAGES - 1;
                if (nk < nkt) load_tile(nk);
                cp_async_commit();
            }
            const double* As = gemm_smem + (size_t)(kt % GEMM_STAGES) * Cfg::STAGE;
            const double* Bs = As + Cfg::A_TILE;
#pragma unroll
            for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
                double af[MI], bf[NJ];
                const int k = kk * 4 + lc;
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int m = wm * WTM + i * 8 + lr;
                    af[i] = (AL == A_MCONTIG) ? As[k * Cfg::PADM + m] : As[m * GEMM_PADK + k];
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int n = wn * WTN + j * 8 + lr;
                    bf[j] = (BL == B_KCONTIG) ? Bs[n * GEMM_PADK + k] : Bs[k * Cfg::PADN + n];
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cp_async_wait<0>();
    }

    // epilogue: thread holds (row = lr, cols = 2*lc + {0,1}) of every 8x8 tile
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m0 + wm * WTM + i * 8 + lr;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn * WTN + j * 8 + 2 * lc + e;
                if (n >= P.N) continue;
                if (P.store_mode == 1 && m < n) continue;
                double v = P.alpha * acc[i][j][e];
                if (!preload && P.beta != 0.0) v += P.beta * P.Cin[(long)m + (long)n * P.ldcin];
                if (P.bias_n) v += P.bias_n[n];
                if (P.C2 && n >= P.nsplit) P.C2[(long)m + (long)(n - P.nsplit) * P.ldc] = v;
                else Cz[(long)m + (long)n * P.ldc] = v;
            }
        }
    }
}

}  // namespace kb
