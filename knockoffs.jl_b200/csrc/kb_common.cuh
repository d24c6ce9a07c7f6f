// Common declarations for libknockoffs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/knockoffs_b200.h"

namespace kb {

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace kb

// The opaque context of the C ABI.  One per host thread / device.
struct kb_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_h2d = nullptr;   // copy streams for the row-chunk pipeline
    cudaStream_t stream_d2h = nullptr;
    cudaStream_t stream_aux = nullptr;   // look-ahead stream of the blocked Cholesky (trailing update || next panel)
    void* green_ctx = nullptr;           // CUgreenCtx that stream_aux belongs to (SM partition), or NULL
    int aux_sms = 0;                     // SMs stream_aux may use (0 = all: no partition)
    int background = 0;                  // kb_create_background: the main stream is confined to the partition as well
    cudaEvent_t aux_ev[2] = {nullptr, nullptr};
    cudaEvent_t ccd_ev[3] = {nullptr, nullptr, nullptr};   // blocked CCD: panel ready / gather done / rank-256 update done
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string last_error;
    unsigned long long launches = 0;     // kernels launched by this library since creation
    double flops = 0.0;                  // FP64 flops issued to the DMMA GEMM kernel since creation (clipped k-ranges counted as clipped)
    long sample_chunk_rows = 0;          // 0 = default (4096)
    int positive_cholesky = 0;           // 1: a failed Cholesky in the conditional-factor stage falls back to the restated
                                         //    cholesky(Positive, ·) (positive.cu, parity unpinned) instead of KB_ERR_NOT_PD
    int positive_modified = 0;           // pivots the last conditional-factor call replaced (0 = ordinary Cholesky factors)
    size_t l2_persist_max = 0;           // cudaDevAttrMaxPersistingL2CacheSize
    size_t l2_window_max = 0;            // cudaDevAttrMaxAccessPolicyWindowSize
    size_t l2_size = 0;
    // workspace pool: device buffers released by a call's Scope are kept for the next call
    // (cudaMalloc / cudaFree synchronise the device and cost ms each at the 200 MB - 2 GB sizes used here)
    struct PoolBlock { void* ptr; size_t bytes; };
    std::vector<PoolBlock> pool_free;
    size_t pool_bytes = 0;
    cudaEvent_t pipe_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // pinned bounce buffers for PAGEABLE host matrices (a Julia Matrix{Float64} / numpy array is pageable): kept between calls,
    // released by kb_trim / kb_destroy.  [0,1] input X chunks, [2,3] input Z chunks, [4,5] output chunks
    void* pinned[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t pinned_cap[6] = {0, 0, 0, 0, 0, 0};
};

namespace kb {

int set_error(kb_ctx* ctx, int code, const char* fmt, ...);

#define KB_CUDA(ctx, expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return kb::set_error((ctx), KB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

#define KB_TRY(expr)                     \
    do {                                 \
        int _rc = (expr);                \
        if (_rc != KB_OK) return _rc;    \
    } while (0)

// scoped device allocations: returned to ctx's workspace pool by Scope's destructor
struct Scope {
    kb_ctx* ctx;
    std::vector<kb_ctx::PoolBlock> blocks;
    explicit Scope(kb_ctx* c) : ctx(c) {}
    ~Scope() {
        for (auto& b : blocks) ctx->pool_free.push_back(b);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        const size_t want = ((count * sizeof(T) + 256 + 255) / 256) * 256;
        // best fit among the free blocks that are not more than 25 % larger than needed
        int best = -1;
        for (int i = 0; i < (int)ctx->pool_free.size(); ++i) {
            const size_t b = ctx->pool_free[i].bytes;
            if (b >= want && b <= want + want / 4 + 4096 && (best < 0 || b < ctx->pool_free[best].bytes)) best = i;
        }
        kb_ctx::PoolBlock blk{nullptr, 0};
        if (best >= 0) {
            blk = ctx->pool_free[best];
            ctx->pool_free.erase(ctx->pool_free.begin() + best);
        } else {
            cudaError_t e = cudaMalloc(&blk.ptr, want);
            if (e != cudaSuccess) {
                // release the cached blocks and retry once
                cudaGetLastError();
                for (auto& f : ctx->pool_free) cudaFree(f.ptr);
                ctx->pool_bytes = 0;
                ctx->pool_free.clear();
                e = cudaMalloc(&blk.ptr, want);
            }
            if (e != cudaSuccess) {
                *out = nullptr;
                cudaGetLastError();
                return set_error(ctx, KB_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            }
            blk.bytes = want;
            ctx->pool_bytes += want;
        }
        blocks.push_back(blk);
        *out = reinterpret_cast<T*>(blk.ptr);
        return KB_OK;
    }
};

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per-function state shared by every host thread / context on the
// device: raise it monotonically under a lock so concurrent contexts (different p, different group sizes) never
// lower each other's limit between the attribute call and the launch.
cudaError_t raise_max_dynamic_smem(const void* func, size_t bytes);

}  // namespace kb
