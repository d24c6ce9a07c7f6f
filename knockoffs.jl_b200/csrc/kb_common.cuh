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
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string last_error;
    std::vector<void*> allocs;           // everything cudaMalloc'ed through ws_alloc (freed per call)
    // pinned staging for host<->device pipelines
    void* pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t pinned_cap[4] = {0, 0, 0, 0};
    unsigned long long launches = 0;     // kernels launched by this library since creation
    long sample_chunk_rows = 0;          // 0 = default (4096)
    size_t l2_persist_max = 0;           // cudaDevAttrMaxPersistingL2CacheSize
    size_t l2_window_max = 0;            // cudaDevAttrMaxAccessPolicyWindowSize
    size_t l2_size = 0;
    cudaEvent_t pipe_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
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

// scoped device allocations: freed by Scope's destructor
struct Scope {
    kb_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scope(kb_ctx* c) : ctx(c) {}
    ~Scope() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T) + 256);
        if (e != cudaSuccess) {
            *out = nullptr;
            return set_error(ctx, KB_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T),
                             cudaGetErrorString(e));
        }
        ptrs.push_back(p);
        *out = reinterpret_cast<T*>(p);
        return KB_OK;
    }
};

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace kb
