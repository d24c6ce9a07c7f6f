// TMA + mbarrier variant of the FP64 tensor-core GEMM (measurement variant, VERDICT r01 item 6 / north_star "TMA-staged
// tiles"): C (M x N) = alpha * A * B + beta * C for the sampling stage's layouts -- A column-major M x K (m contiguous),
// B column-major K x N (k contiguous) -- with the operand tiles staged by `cp.async.bulk.tensor` (SASS UTMALDG) into
// 128-byte-swizzled shared memory and a 4-stage full / empty mbarrier ring fed by ONE producer warp, instead of the
// LDGSTS (cp.async) pipeline + padded rows of dgemm_dmma.cuh.  The math is the same warp-level DMMA (mma.sync m8n8k4.f64;
// tcgen05 has no f64 kind).
//
// Shared-memory layout per stage (16 KB, no padding):
//   A: BM/16 boxes of {16 m, 16 k}: k-row of 128 bytes, 16-byte chunks XOR-swizzled by (k & 7)     (SWIZZLE_128B)
//   B: one box {16 k, BN n}:        n-row of 128 bytes, 16-byte chunks XOR-swizzled by (n & 7)
// A DMMA fragment reads (slot = lane / 4, k = lane % 4).  With the identity slot -> row map the swizzle would leave 2-way
// bank conflicts, so the 8 slots of a fragment are mapped to rows by a fixed permutation (sigma for A inside a 16-row box,
// rho for B inside an 8-row group) under which the 16 lanes of a half-warp hit 16 distinct 8-byte banks; the epilogue
// applies the same permutation to the output coordinates.
#include "kb_common.cuh"
#include "dgemm_dmma.cuh"
#include "pipeline.cuh"
#include <cuda.h>
#include <cstdlib>

namespace kb {

constexpr int TBM = 64, TBN = 64, TBK = 16, TSTAGES = 4;
constexpr int T_CONSUMER_WARPS = 4, T_THREADS = 32 * (T_CONSUMER_WARPS + 1);
constexpr int T_A_BYTES = TBM * TBK * 8, T_B_BYTES = TBN * TBK * 8, T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
constexpr size_t T_SMEM = (size_t)TSTAGES * T_STAGE_BYTES + 1024 /* alignment slack */ + 2 * TSTAGES * 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// slot -> row permutations (see the header comment)
__device__ __forceinline__ int sigma16(int slot, int half) { return (slot & 1) + ((slot >> 1) & 1) * 8 + (slot >> 2) * 2 + half * 4; }
__device__ __forceinline__ int rho8(int slot) { return (slot & 3) * 2 + (slot >> 2); }

__global__ void __launch_bounds__(T_THREADS, 3)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int N, int K,
                 double alpha, double beta, double* __restrict__ C, long ldc) {
    extern __shared__ unsigned char tsm_raw[];
    unsigned char* tsm = (unsigned char*)(((uintptr_t)tsm_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B: 1024-byte aligned
    unsigned long long* full_bar = (unsigned long long*)(tsm + (size_t)TSTAGES * T_STAGE_BYTES);
    unsigned long long* empty_bar = full_bar + TSTAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // same grouped rasterisation as dgemm_dmma_kernel
    int bx, by;
    {
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
    const int m0 = bx * TBM, n0 = by * TBN;
    const int nkt = (K + TBK - 1) / TBK;
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) {
            mbar_init(&full_bar[s], 1);                       // the producer's arrive.expect_tx (+ the TMA byte count)
            mbar_init(&empty_bar[s], T_CONSUMER_WARPS);       // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == T_CONSUMER_WARPS) {
        // ---- producer warp: one lane feeds the ring ----
        if (lane == 0) {
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % TSTAGES;
                const unsigned round = (unsigned)(kt / TSTAGES);
                if (kt >= TSTAGES) mbar_wait(&empty_bar[s], (round - 1) & 1);     // all consumer warps released use `round − 1`
                unsigned char* As = tsm + (size_t)s * T_STAGE_BYTES;
                unsigned char* Bs = As + T_A_BYTES;
                mbar_expect_tx(&full_bar[s], T_STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < TBM / 16; ++j) tma_load_2d(As + j * 2048, &mapA, m0 + 16 * j, kt * TBK, &full_bar[s]);
                tma_load_2d(Bs, &mapB, kt * TBK, n0, &full_bar[s]);
            }
        }
        return;
    }
    // ---- consumer warps: 2 x 2 layout, 32 x 32 per warp = 4 x 4 DMMA tiles ----
    const int wm = warp >> 1, wn = warp & 1;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // per-thread constant parts of the fragment addresses (in doubles)
    int a_off[4], a_mm[4], b_off[4], b_n7[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int box = wm * 2 + (i >> 1);
        const int mm = sigma16(lr, i & 1);
        a_off[i] = box * 256 + (mm & 1);
        a_mm[i] = mm >> 1;                       // 16-byte chunk of this row inside the 128-byte k-row
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = wn * 32 + j * 8 + rho8(lr);
        b_off[j] = n * 16;
        b_n7[j] = n & 7;
    }
    for (int kt = 0; kt < nkt; ++kt) {
        const int s = kt % TSTAGES;
        mbar_wait(&full_bar[s], (unsigned)(kt / TSTAGES) & 1);
        const double* As = (const double*)(tsm + (size_t)s * T_STAGE_BYTES);
        const double* Bs = As + T_A_BYTES / 8;
#pragma unroll
        for (int kk = 0; kk < TBK / 4; ++kk) {
            const int k = kk * 4 + lc;
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[a_off[i] + k * 16 + ((a_mm[i] ^ (k & 7)) << 1)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[b_off[j] + (((k >> 1) ^ b_n7[j]) << 1) + (k & 1)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // ---- epilogue (permuted coordinates) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + (i >> 1) * 16 + sigma16(lr, i & 1);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn * 32 + j * 8 + rho8(2 * lc + e);
                if (n >= N) continue;
                double v = alpha * acc[i][j][e];
                if (beta != 0.0) v += beta * C[(long)m + (long)n * ldc];
                C[(long)m + (long)n * ldc] = v;
            }
        }
    }
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// The sampling stage's batched launch (pipeline.cu sample_chunk, structured form) on the TMA pipeline:
//   C_z (M x N) = Cin + A0_z · B0_z [+ A1_z · B1_z],   z = blockIdx.z < batch
// with B0 lower triangular in (k, n) (k >= n0) and B1 upper triangular up to a band (k < n0 + BN + slack); the last problem
// has one segment only.  A_s: (m, k, z) tensor maps, B_s: (k, n, z) tensor maps.  Same ring, swizzle and fragment
// permutations as dgemm_tma_kernel; the k-tile counter runs on across the two segments, so the ring never drains.
struct SampleTmaParams {
    int M, N, K;
    int batch, last_nseg, nseg;
    int khi_slack;
    const double* Cin;
    long ldcin;
    double* C;
    long ldc, zC;
};

__global__ void __launch_bounds__(T_THREADS, 3)
sample_tma_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
                  const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1, const SampleTmaParams P) {
    extern __shared__ unsigned char tsm_raw[];
    unsigned char* tsm = (unsigned char*)(((uintptr_t)tsm_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* full_bar = (unsigned long long*)(tsm + (size_t)TSTAGES * T_STAGE_BYTES);
    unsigned long long* empty_bar = full_bar + TSTAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bx, by;
    {
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
    const int m0 = bx * TBM, n0 = by * TBN;
    const int z = blockIdx.z;
    const int nseg = (P.batch > 1 && P.last_nseg > 0 && z == P.batch - 1) ? P.last_nseg : P.nseg;
    // k-tile ranges of the two segments (tile aligned; the operands are exactly zero outside the clipped ranges)
    int kb[2], nk[2];
    {
        int kbeg = (n0 / TBK) * TBK, kend = P.K;                                 // segment 0: SEG_KLO_N
        kb[0] = kbeg; nk[0] = (kend > kbeg) ? (kend - kbeg + TBK - 1) / TBK : 0;
        kbeg = 0; kend = min(P.K, n0 + TBN + P.khi_slack);                       // segment 1: SEG_KHI_N
        kb[1] = 0; nk[1] = (nseg > 1 && kend > 0) ? (kend + TBK - 1) / TBK : 0;
    }
    const int ntot = nk[0] + nk[1];
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], T_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == T_CONSUMER_WARPS) {
        if (lane == 0) {
            for (int t = 0; t < ntot; ++t) {
                const int s = t % TSTAGES;
                const unsigned round = (unsigned)(t / TSTAGES);
                if (t >= TSTAGES) mbar_wait(&empty_bar[s], (round - 1) & 1);
                unsigned char* As = tsm + (size_t)s * T_STAGE_BYTES;
                unsigned char* Bs = As + T_A_BYTES;
                const int seg = (t < nk[0]) ? 0 : 1;
                const int k0 = kb[seg] + (seg == 0 ? t : t - nk[0]) * TBK;
                const CUtensorMap* ma = seg == 0 ? &mapA0 : &mapA1;
                const CUtensorMap* mb = seg == 0 ? &mapB0 : &mapB1;
                mbar_expect_tx(&full_bar[s], T_STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < TBM / 16; ++j) tma_load_3d(As + j * 2048, ma, m0 + 16 * j, k0, z, &full_bar[s]);
                tma_load_3d(Bs, mb, k0, n0, z, &full_bar[s]);
            }
        }
        return;
    }
    const int wm = warp >> 1, wn = warp & 1;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[4][4][2];
    // Cin enters through the accumulators (alpha = beta = 1), permuted coordinates
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + (i >> 1) * 16 + sigma16(lr, i & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn * 32 + j * 8 + rho8(2 * lc + e);
                acc[i][j][e] = (m < P.M && n < P.N) ? P.Cin[(long)m + (long)n * P.ldcin] : 0.0;
            }
    }
    int a_off[4], a_mm[4], b_off[4], b_n7[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int box = wm * 2 + (i >> 1);
        const int mm = sigma16(lr, i & 1);
        a_off[i] = box * 256 + (mm & 1);
        a_mm[i] = mm >> 1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = wn * 32 + j * 8 + rho8(lr);
        b_off[j] = n * 16;
        b_n7[j] = n & 7;
    }
    for (int t = 0; t < ntot; ++t) {
        const int s = t % TSTAGES;
        mbar_wait(&full_bar[s], (unsigned)(t / TSTAGES) & 1);
        const double* As = (const double*)(tsm + (size_t)s * T_STAGE_BYTES);
        const double* Bs = As + T_A_BYTES / 8;
#pragma unroll
        for (int kk = 0; kk < TBK / 4; ++kk) {
            const int k = kk * 4 + lc;
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[a_off[i] + k * 16 + ((a_mm[i] ^ (k & 7)) << 1)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[b_off[j] + (((k >> 1) ^ b_n7[j]) << 1) + (k & 1)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    double* const Cz = P.C + (long)z * P.zC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + (i >> 1) * 16 + sigma16(lr, i & 1);
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn * 32 + j * 8 + rho8(2 * lc + e);
                if (n < P.N) Cz[(long)m + (long)n * P.ldc] = acc[i][j][e];
            }
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*fn_encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static fn_encode_tiled get_encode() {
    static fn_encode_tiled f = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (fn_encode_tiled)p;
    }();
    return f;
}

// column-major rows x cols matrix of doubles (leading dimension ld), box {b0 rows, b1 cols}, 128-byte swizzle
static int make_map(kb_ctx* ctx, CUtensorMap* map, const double* base, long rows, long cols, long ld, int b0, int b1) {
    fn_encode_tiled enc = get_encode();
    if (!enc) return set_error(ctx, KB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ctx, KB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return KB_OK;
}

// rows x cols x batch (column-major slices of leading dimension ld, `zstride` doubles apart), box {b0, b1, 1}
static int make_map3(kb_ctx* ctx, CUtensorMap* map, const double* base, long rows, long cols, long ld, long batch, long zstride,
                     int b0, int b1) {
    fn_encode_tiled enc = get_encode();
    if (!enc) return set_error(ctx, KB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)(batch < 1 ? 1 : batch)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(double), (cuuint64_t)(zstride > 0 ? zstride : ld * cols) * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ctx, KB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return KB_OK;
}

bool sample_tma_usable(const double* T, long ldt, long tstride, const double* L, long ldl, long lstride, const double* Nn, long ldn,
                       long nstride) {
    // OFF by default: measured in the sampling stage at p = 5000, m = 5 (r02j) the TMA kernel is 0.6 % slower than the LDGSTS
    // kernel (1685.4 vs 1675.6 ms per step) although it is 1.4-2.2 % faster on plain GEMM shapes (profiles/r02_gemm_variants.md).
    // KB_SAMPLE_TMA=1 switches it on (read at every plan creation, so a test can toggle it).
    const char* e = getenv("KB_SAMPLE_TMA");
    const bool on = e && e[0] == '1';
    auto ok = [](const void* p_, long ld, long zs) { return !((uintptr_t)p_ & 15) && !(ld & 1) && !(zs & 1); };
    return on && get_encode() && ok(T, ldt, tstride) && ok(L, ldl, lstride) && (Nn == nullptr || ok(Nn, ldn, nstride));
}

// The 4 tensor maps of the structured sampling launch (constant for a SamplePlan): A0 = S_k (T, z-stride tstride),
// B0 = L_kk (z-stride 2 bstride), A1 = S_{k+1} (T + tstride), B1 = N_k.
int sample_tma_maps(kb_ctx* ctx, SampleTmaMaps* out, const double* T, long chunk, int p, int m, long tstride, const double* Lb, long ldo,
                    long bstride, const double* Nn, long ldn) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(out->raw);
    KB_TRY(make_map3(ctx, &maps[0], T, chunk, p, chunk, m, tstride, 16, TBK));
    KB_TRY(make_map3(ctx, &maps[1], Lb, p, p, ldo, m, 2 * bstride, TBK, TBN));
    if (m > 1) {
        KB_TRY(make_map3(ctx, &maps[2], T + tstride, chunk, p, chunk, m - 1, tstride, 16, TBK));
        KB_TRY(make_map3(ctx, &maps[3], Nn, p, p, ldn, m - 1, ldn * (long)p, TBK, TBN));
    } else {
        maps[2] = maps[0];
        maps[3] = maps[1];
    }
    return KB_OK;
}

// X̃_z = Mu + S_z L_zz + S_{z+1} N_z for z < m (the last copy has no N block): ONE launch, gridDim.z = m.
int sample_tma_launch(kb_ctx* ctx, const SampleTmaMaps& mp, int rows, int p, int m, int band, const double* Mu, long ldmu, double* Xko,
                      long ldxko) {
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(mp.raw);
    SampleTmaParams P{};
    P.M = rows; P.N = p; P.K = p; P.batch = m; P.nseg = (m > 1) ? 2 : 1; P.last_nseg = 1; P.khi_slack = band;
    P.Cin = Mu; P.ldcin = ldmu; P.C = Xko; P.ldc = ldxko; P.zC = (long)p * ldxko;
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)sample_tma_kernel, T_SMEM));
    dim3 grid(ceil_div(rows, TBM), ceil_div(p, TBN), m);
    sample_tma_kernel<<<grid, T_THREADS, T_SMEM, ctx->stream>>>(maps[0], maps[1], maps[2], maps[3], P);
    ctx->launches++;
    ctx->flops += (2.0 * m - 1.0) * (double)rows * (double)p * (double)p;      // two triangular products per copy, one for the last
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

// C = alpha * A * B + beta * C, A: M x K (m contiguous, lda), B: K x N (k contiguous, ldb); device pointers; enqueued on
// ctx's stream.  Base pointers 16-byte aligned and leading dimensions even (TMA global strides are multiples of 16 bytes).
int gemm_tma_nn(kb_ctx* ctx, int M, int N, int K, double alpha, const double* A, long lda, const double* B, long ldb, double beta,
                double* C, long ldc) {
    if (M <= 0 || N <= 0 || K <= 0) return KB_OK;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1))
        return set_error(ctx, KB_ERR_INVALID, "gemm_tma: operands must be 16-byte aligned with even leading dimensions");
    CUtensorMap mapA, mapB;
    KB_TRY(make_map(ctx, &mapA, A, M, K, lda, 16, TBK));       // dim0 = m (contiguous), dim1 = k
    KB_TRY(make_map(ctx, &mapB, B, K, N, ldb, TBK, TBN));      // dim0 = k (contiguous), dim1 = n
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)dgemm_tma_kernel, T_SMEM));
    dim3 grid(ceil_div(M, TBM), ceil_div(N, TBN));
    dgemm_tma_kernel<<<grid, T_THREADS, T_SMEM, ctx->stream>>>(mapA, mapB, M, N, K, alpha, beta, C, ldc);
    ctx->launches++;
    ctx->flops += 2.0 * M * N * (double)K;
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace kb

extern "C" int kb_test_dgemm_tma_dev(kb_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* dA, int64_t lda,
                                     const double* dB, int64_t ldb, double beta, double* dC, int64_t ldc) {
    if (!ctx) return KB_ERR_INVALID;
    if (!dA || !dB || !dC || M < 0 || N < 0 || K < 0) return kb::set_error(ctx, KB_ERR_INVALID, "test_dgemm_tma: bad arguments");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != ctx->device) cudaSetDevice(ctx->device);
    const int rc = kb::gemm_tma_nn(ctx, (int)M, (int)N, (int)K, alpha, dA, (long)lda, dB, (long)ldb, beta, dC, (long)ldc);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    return rc;
}
