// extern "C" entry points of libknockoffs_b200 (see include/knockoffs_b200.h for the contract and the
// reference lines each call replaces).  Host-pointer calls copy in/out; nothing is retained after return.
#include "kb_common.cuh"
#include "linalg.cuh"
#include "pipeline.cuh"
#include "group.cuh"
#include "stats.cuh"
#include "approx.cuh"
#include "rep.cuh"
#include "hostpipe.cuh"
#include "positive.cuh"
#include <cuda.h>
#include <algorithm>
#include <unordered_map>
#include <cstring>
#include <cstdlib>
#include <new>
#include <thread>

using namespace kb;

namespace kb {

// Marginal-statistic accumulators fed by the sampling pipeline chunk by chunk (fit_marginal: X̃ is consumed
// while it is still on the device).

bool host_is_pageable(const void* ptr) {
    if (!ptr) return false;
    static const bool enabled = [] { const char* e = getenv("KB_PAGEABLE_STAGING"); return !(e && e[0] == '0'); }();
    if (!enabled) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int ensure_pinned(kb_ctx* ctx, int idx, size_t bytes, double** out) {
    if (ctx->pinned_cap[idx] < bytes) {
        if (ctx->pinned[idx]) cudaFreeHost(ctx->pinned[idx]);
        ctx->pinned[idx] = nullptr;
        ctx->pinned_cap[idx] = 0;
        KB_CUDA(ctx, cudaHostAlloc(&ctx->pinned[idx], bytes, cudaHostAllocDefault));
        ctx->pinned_cap[idx] = bytes;
    }
    *out = static_cast<double*>(ctx->pinned[idx]);
    return KB_OK;
}

void release_pinned(kb_ctx* ctx) {
    for (int i = 0; i < 6; ++i) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        ctx->pinned[i] = nullptr;
        ctx->pinned_cap[i] = 0;
    }
}

void parallel_copy2d(double* dst, long dld, const double* src, long sld, long rows, long cols) {
    static const int nthreads = [] {
        const char* e = getenv("KB_HOST_COPY_THREADS");
        int t = e ? atoi(e) : (int)std::thread::hardware_concurrency() / 4;
        return t < 1 ? 1 : (t > 16 ? 16 : t);
    }();
    auto work = [=](long c0, long c1) {
        for (long c = c0; c < c1; ++c) memcpy(dst + c * dld, src + c * sld, sizeof(double) * (size_t)rows);
    };
    const int nt = (int)std::min<long>(nthreads, std::max<long>(1, (rows * cols) >> 16));
    if (nt <= 1) { work(0, cols); return; }
    std::vector<std::thread> th;
    const long per = (cols + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, std::min(cols, per * t), std::min(cols, per * (t + 1)));
    work(0, std::min(cols, per));
    for (auto& x : th) x.join();
}

// Joins the scatter threads of the pageable-output path however the pipeline function is left.
struct ThreadJoiner {
    std::thread t[2];
    ~ThreadJoiner() {
        for (auto& x : t)
            if (x.joinable()) x.join();
    }
};

// Row-chunk pipeline for host-resident X / Z / X̃: H2D of chunk i+1 and D2H of chunk i-1 overlap the
// GEMMs of chunk i (three streams, double-buffered device slabs).  Xko == NULL: X̃ is not copied back.
int sample_host_pipeline(kb_ctx* ctx, const double* X, long n, int p, const double* dmu, const double* dSiS,
                         const double* dLb, long ldo, int m, const double* Z, uint64_t seed, int64_t row_offset,
                         double* Xko, const StatsHook* hook, long hld) {
    if (n <= 0) return KB_OK;
    if (hld < n) hld = n;            // leading dimension of the HOST matrices X, Z, Xko (> n: a row shard of larger matrices)
    long chunk = ctx->sample_chunk_rows > 0 ? ctx->sample_chunk_rows : 4096;
    if (chunk > n) chunk = n;
    chunk = (chunk + 1) & ~1L;
    Scope sc(ctx);
    StreamDrain drain(ctx);          // error paths: nothing may still be in flight when the slabs go back to the pool
    SamplePlan plan;
    KB_TRY(sample_plan_create(ctx, sc, p, m, dmu, dSiS, dLb, ldo, chunk, Z == nullptr, &plan));
    double *dXc[2], *dZc[2] = {nullptr, nullptr}, *dOut[2];
    for (int b = 0; b < 2; ++b) {
        KB_TRY(sc.alloc(&dXc[b], (size_t)chunk * p));
        if (Z) KB_TRY(sc.alloc(&dZc[b], (size_t)chunk * p * m));
        KB_TRY(sc.alloc(&dOut[b], (size_t)chunk * p * m));
    }
    // PAGEABLE host matrices (what a Julia or numpy caller passes): cudaMemcpy2DAsync from / to pageable memory is staged
    // by the driver synchronously and chunk by chunk.  Instead the rows of a chunk are gathered into / scattered from pinned
    // bounce buffers by a few host threads (32 KB contiguous per column), so the copies stay asynchronous and overlap the GEMMs.
    const bool x_page = host_is_pageable(X), z_page = Z && host_is_pageable(Z), o_page = Xko && host_is_pageable(Xko);
    double *sX[2] = {nullptr, nullptr}, *sZ[2] = {nullptr, nullptr}, *sO[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        if (x_page) KB_TRY(ensure_pinned(ctx, b, sizeof(double) * (size_t)chunk * p, &sX[b]));
        if (z_page) KB_TRY(ensure_pinned(ctx, 2 + b, sizeof(double) * (size_t)chunk * p * m, &sZ[b]));
        if (o_page) KB_TRY(ensure_pinned(ctx, 4 + b, sizeof(double) * (size_t)chunk * p * m, &sO[b]));
    }
    ThreadJoiner scatter;                      // declared after `drain`: joined first, then the streams are drained
    cudaEvent_t* ev_h2d = ctx->pipe_ev;        // [2]
    cudaEvent_t* ev_comp = ctx->pipe_ev + 2;   // [2]
    cudaEvent_t* ev_d2h = ctx->pipe_ev + 4;    // [2]
    const int device = ctx->device;
    // the plan's constants were enqueued on ctx->stream; the copy streams only touch their own slabs
    long i = 0;
    for (long r0 = 0; r0 < n; r0 += chunk, ++i) {
        const int b = (int)(i & 1);
        const long rows = std::min(chunk, n - r0);
        if (i >= 2) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ev_comp[b], 0));
        if ((x_page || z_page) && i >= 2) KB_CUDA(ctx, cudaEventSynchronize(ev_h2d[b]));   // bounce buffers b are free again
        if (x_page) {
            parallel_copy2d(sX[b], chunk, X + r0, hld, rows, p);
            KB_TRY(h2d_matrix(ctx, sX[b], chunk, dXc[b], chunk, rows, p, ctx->stream_h2d));
        } else {
            KB_TRY(h2d_matrix(ctx, X + r0, hld, dXc[b], chunk, rows, p, ctx->stream_h2d));
        }
        if (Z) {
            if (z_page) {
                parallel_copy2d(sZ[b], chunk, Z + r0, hld, rows, (long)p * m);
                KB_TRY(h2d_matrix(ctx, sZ[b], chunk, dZc[b], chunk, rows, (long)p * m, ctx->stream_h2d));
            } else {
                KB_TRY(h2d_matrix(ctx, Z + r0, hld, dZc[b], chunk, rows, (long)p * m, ctx->stream_h2d));
            }
        }
        KB_CUDA(ctx, cudaEventRecord(ev_h2d[b], ctx->stream_h2d));
        KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[b], 0));
        if (i >= 2) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_d2h[b], 0));
        KB_TRY(sample_chunk(ctx, plan, dXc[b], rows, chunk, Z ? dZc[b] : nullptr, chunk, seed, row_offset + r0, dOut[b],
                            chunk));
        if (hook) {
            KB_TRY(marginal_accumulate_dev(ctx, dXc[b], rows, chunk, p, hook->dys + r0, hook->dacc_x, r0 == 0));
            KB_TRY(marginal_accumulate_dev(ctx, dOut[b], rows, chunk, (long)p * m, hook->dys + r0, hook->dacc_ko, r0 == 0));
        }
        KB_CUDA(ctx, cudaEventRecord(ev_comp[b], ctx->stream));
        KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_d2h, ev_comp[b], 0));
        if (Xko && o_page) {
            if (scatter.t[b].joinable()) scatter.t[b].join();                 // chunk i − 2 has left bounce buffer b
            KB_TRY(d2h_matrix(ctx, dOut[b], chunk, sO[b], chunk, rows, (long)p * m, ctx->stream_d2h));
            KB_CUDA(ctx, cudaEventRecord(ev_d2h[b], ctx->stream_d2h));
            cudaEvent_t evb = ev_d2h[b];
            double* src = sO[b];
            double* dst = Xko + r0;
            const long cols = (long)p * m;
            scatter.t[b] = std::thread([=] {
                cudaSetDevice(device);
                cudaEventSynchronize(evb);
                parallel_copy2d(dst, hld, src, chunk, rows, cols);
            });
        } else {
            if (Xko) KB_TRY(d2h_matrix(ctx, dOut[b], chunk, Xko + r0, hld, rows, (long)p * m, ctx->stream_d2h));
            KB_CUDA(ctx, cudaEventRecord(ev_d2h[b], ctx->stream_d2h));
        }
    }
    for (auto& t : scatter.t)
        if (t.joinable()) t.join();
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_h2d));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_d2h));
    return KB_OK;
}

// G (ld x p, lower tiles) += X_rows' X_rows and colsum += 1' X_rows for `n` rows of a HOST matrix (leading dimension hld),
// streamed in row chunks: H2D of chunk i + 1 overlaps the DMMA SYRK of chunk i.  dG / dcs must be zeroed by the caller.
int gram_accumulate_host_rows(kb_ctx* ctx, const double* X, long n, long hld, int p, double* dG, long ld, double* dcs) {
    if (n <= 0) return KB_OK;
    Scope sc(ctx);
    StreamDrain drain(ctx);
    const bool x_page = host_is_pageable(X);
    long chunk = x_page ? 16384 : 65536;       // pageable input goes through pinned bounce buffers: smaller chunks
    if (chunk > n) chunk = (n + 1) & ~1L;
    double *dXc[2], *sX[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        KB_TRY(sc.alloc(&dXc[b], (size_t)chunk * p));
        if (x_page) KB_TRY(ensure_pinned(ctx, b, sizeof(double) * (size_t)chunk * p, &sX[b]));
    }
    cudaEvent_t* ev_h2d = ctx->pipe_ev;
    cudaEvent_t* ev_comp = ctx->pipe_ev + 2;
    long i = 0;
    for (long r0 = 0; r0 < n; r0 += chunk, ++i) {
        const int b = (int)(i & 1);
        const long rows = std::min(chunk, n - r0);
        if (i >= 2) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ev_comp[b], 0));
        if (x_page) {
            if (i >= 2) KB_CUDA(ctx, cudaEventSynchronize(ev_h2d[b]));
            parallel_copy2d(sX[b], chunk, X + r0, hld, rows, p);
            KB_TRY(h2d_matrix(ctx, sX[b], chunk, dXc[b], chunk, rows, p, ctx->stream_h2d));
        } else {
            KB_TRY(h2d_matrix(ctx, X + r0, hld, dXc[b], chunk, rows, p, ctx->stream_h2d));
        }
        KB_CUDA(ctx, cudaEventRecord(ev_h2d[b], ctx->stream_h2d));
        KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[b], 0));
        KB_TRY(gram_partial_dev(ctx, dXc[b], rows, chunk, p, dG, ld, dcs, 1));
        KB_CUDA(ctx, cudaEventRecord(ev_comp[b], ctx->stream));
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_h2d));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// condition(X, μ, Σ, S; m) (modelX.jl:96-122) on host matrices; X / Z / Xko are n rows of matrices with leading dimension hld.
int condition_host(kb_ctx* ctx, const double* X, long n, long hld, int p, const double* mu, const double* Sigma, const double* S,
                   int s_is_diag, int m, const double* Z, uint64_t seed, int64_t row_offset, double* Xko) {
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);   // modelX.jl:105
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dSig, *dS, *dmu, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&dS, s_is_diag ? (size_t)p : pp));
    KB_TRY(sc.alloc(&dmu, p));
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_CUDA(ctx, cudaMemcpyAsync(dSig, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, S, sizeof(double) * (s_is_diag ? (size_t)p : pp), cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(cond_factor_dev(ctx, dSig, (long)p, p, dS, (long)p, s_is_diag, m, dSiS, dLb, (long)p));
    return sample_host_pipeline(ctx, X, n, p, dmu, dSiS, dLb, (long)p, m, Z, seed, row_offset, Xko, nullptr, hld);
}

}  // namespace kb

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int bad_ctx() { return KB_ERR_INVALID; }

// SM partition for the auxiliary stream (CUDA green context, driver API resolved through the runtime -- no libcuda link):
// stream_aux may use all SMs but `reserve` of them, so the single-CTA chain kernels of the main stream (blocked CCD
// sequential kernel: 256 threads x 192 registers, i.e. it needs an SM that holds NO tensor-core GEMM CTA; Cholesky diagonal
// kernel) always find a free SM while a rank-256 pass / trailing update runs beside them.  Without the partition the block
// scheduler back-fills every freed slot with GEMM CTAs and the chain kernel waits for the GEMM grid to drain.
// Returns false (and leaves ctx untouched) if the driver refuses any step; the caller then creates an ordinary stream.
bool make_green_aux_stream(kb_ctx* ctx, int reserve, int priority, bool also_main = false) {
    auto get = [](const char* name) -> void* {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return f;
    };
    typedef CUresult (*fn_devget)(CUdevice*, int);
    typedef CUresult (*fn_getres)(CUdevice, CUdevResource*, CUdevResourceType);
    typedef CUresult (*fn_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    typedef CUresult (*fn_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    typedef CUresult (*fn_gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*fn_gstream)(CUstream*, CUgreenCtx, unsigned int, int);
    typedef CUresult (*fn_gdestroy)(CUgreenCtx);
    auto devget = (fn_devget)get("cuDeviceGet");
    auto getres = (fn_getres)get("cuDeviceGetDevResource");
    auto split = (fn_split)get("cuDevSmResourceSplitByCount");
    auto mkdesc = (fn_desc)get("cuDevResourceGenerateDesc");
    auto gcreate = (fn_gcreate)get("cuGreenCtxCreate");
    auto gstream = (fn_gstream)get("cuGreenCtxStreamCreate");
    auto gdestroy = (fn_gdestroy)get("cuGreenCtxDestroy");
    if (!devget || !getres || !split || !mkdesc || !gcreate || !gstream || !gdestroy) return false;
    CUdevice dev;
    if (devget(&dev, ctx->device) != CUDA_SUCCESS) return false;
    CUdevResource all, grp[1], rest;
    if (getres(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int ngroups = 1;
    if (split(grp, &ngroups, &all, &rest, 0, (unsigned)reserve) != CUDA_SUCCESS || ngroups != 1) return false;
    if (rest.sm.smCount < 32) return false;
    CUdevResourceDesc desc;
    if (mkdesc(&desc, &rest, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx g;
    if (gcreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream st;
    if (gstream(&st, g, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) { gdestroy(g); return false; }
    // events of the primary context must order work across the two kinds of streams: check it once, here
    cudaEvent_t e = nullptr;
    bool ok = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventRecord(e, (cudaStream_t)st) == cudaSuccess;
    ok = ok && cudaStreamWaitEvent(ctx->stream, e, 0) == cudaSuccess;
    ok = ok && cudaEventRecord(e, ctx->stream) == cudaSuccess;
    ok = ok && cudaStreamWaitEvent((cudaStream_t)st, e, 0) == cudaSuccess;
    ok = ok && cudaStreamSynchronize((cudaStream_t)st) == cudaSuccess && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (e) cudaEventDestroy(e);
    if (!ok) {
        cudaGetLastError();
        cudaStreamDestroy((cudaStream_t)st);
        gdestroy(g);
        return false;
    }
    if (also_main) {
        // background context: the MAIN stream lives in the partition too (a second stream of the same green context)
        CUstream st2;
        if (gstream(&st2, g, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) {
            cudaStreamDestroy((cudaStream_t)st);
            gdestroy(g);
            return false;
        }
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        ctx->stream = (cudaStream_t)st2;
    }
    ctx->stream_aux = (cudaStream_t)st;
    ctx->green_ctx = (void*)g;
    ctx->aux_sms = (int)rest.sm.smCount;
    return true;
}
void destroy_green(kb_ctx* ctx) {
    if (!ctx->green_ctx) return;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &f, cudaEnableDefault, &st) == cudaSuccess && f)
        ((CUresult(*)(CUgreenCtx))f)((CUgreenCtx)ctx->green_ctx);
    ctx->green_ctx = nullptr;
}

// Accumulate the marginal-statistic sums of a HOST matrix A (n x ncols, ld n) against dys: row chunks, H2D of
// chunk i+1 overlapping the reduction of chunk i.
int host_marginal_accumulate(kb_ctx* ctx, const double* A, long n, long ncols, const double* dys, double* dacc) {
    Scope sc(ctx);
    StreamDrain drain(ctx);
    long chunk = (long)(((size_t)256 << 20) / (sizeof(double) * (size_t)std::max(1L, ncols)));   // ~256 MB slabs
    chunk = std::max(1024L, std::min(chunk, n));
    if (chunk > n) chunk = n;
    double* dA[2];
    for (int b = 0; b < 2; ++b) KB_TRY(sc.alloc(&dA[b], (size_t)chunk * ncols));
    cudaEvent_t* ev_h2d = ctx->pipe_ev;
    cudaEvent_t* ev_comp = ctx->pipe_ev + 2;
    long i = 0;
    for (long r0 = 0; r0 < n; r0 += chunk, ++i) {
        const int b = (int)(i & 1);
        const long rows = std::min(chunk, n - r0);
        if (i >= 2) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ev_comp[b], 0));
        KB_TRY(h2d_matrix(ctx, A + r0, n, dA[b], chunk, rows, ncols, ctx->stream_h2d));
        KB_CUDA(ctx, cudaEventRecord(ev_h2d[b], ctx->stream_h2d));
        KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[b], 0));
        KB_TRY(marginal_accumulate_dev(ctx, dA[b], rows, chunk, ncols, dys + r0, dacc, r0 == 0));
        KB_CUDA(ctx, cudaEventRecord(ev_comp[b], ctx->stream));
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_h2d));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// fit_marginal's tail (fit_lasso.jl:223-256) from the importance scores dT = [T0 | T1 | ... | Tm] ((m+1) p, device):
// optional group sums, MK_statistics, one threshold per target FDR, the selected sets.  Host outputs; optional ones
// may be NULL.  selected_out: n_w x nfdr 0/1 mask (column f = findall(W .>= τ̂_f)).
int knockoff_filter_from_T(kb_ctx* ctx, const double* dT, long p, int m, const int64_t* groups, const double* fdrs, int nfdr,
                           int knockoff_plus, long rej_bounds, double* W_out, int64_t* kappa_out, double* tau_out,
                           double* taus_out, int32_t* selected_out, int64_t* n_w_out) {
    if (nfdr < 0 || (nfdr > 0 && (!fdrs || !taus_out))) return set_error(ctx, KB_ERR_INVALID, "knockoff filter: bad fdr arguments");
    if (m > 1 && !knockoff_plus)                                   // threshold.jl:59
        return set_error(ctx, KB_ERR_INVALID, "Multiple knockoffs needs to use :knockoff_plus filtering method");
    Scope sc(ctx);
    long nw = p;
    const double* dTs = dT;
    if (groups) {
        // unique(groups) in order of first appearance (fit_lasso.jl:225)
        std::unordered_map<int64_t, int> ids;
        std::vector<int> gidx((size_t)p);
        for (long j = 0; j < p; ++j) {
            auto it = ids.find(groups[j]);
            if (it == ids.end()) it = ids.emplace(groups[j], (int)ids.size()).first;
            gidx[(size_t)j] = it->second;
        }
        nw = (long)ids.size();
        int* dgidx;
        double* dTg;
        KB_TRY(sc.alloc(&dgidx, (size_t)p));
        KB_TRY(sc.alloc(&dTg, (size_t)(m + 1) * nw));
        KB_CUDA(ctx, cudaMemcpyAsync(dgidx, gidx.data(), sizeof(int) * p, cudaMemcpyHostToDevice, ctx->stream));
        KB_TRY(group_abs_sum_dev(ctx, dT, p, m + 1, dgidx, (int)nw, dTg));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // gidx goes out of scope
        dTs = dTg;
    }
    double *dW, *dtau, *dth;
    long long* dkap;
    int* dsel;
    KB_TRY(sc.alloc(&dW, (size_t)nw));
    KB_TRY(sc.alloc(&dtau, (size_t)nw));
    KB_TRY(sc.alloc(&dkap, (size_t)nw));
    KB_TRY(sc.alloc(&dth, (size_t)std::max(1, nfdr)));
    KB_TRY(sc.alloc(&dsel, (size_t)nw * std::max(1, nfdr)));
    KB_TRY(mk_statistics_dev(ctx, dTs, dTs + nw, nw, m, dkap, dtau, dW));
    for (int f = 0; f < nfdr; ++f) {
        if (m > 1) KB_TRY(mk_threshold_dev(ctx, dtau, dkap, nw, m, fdrs[f], rej_bounds, dth + f));   // fit_lasso.jl:249
        else KB_TRY(threshold_dev(ctx, dW, nw, fdrs[f], knockoff_plus ? 1 : 0, rej_bounds, dth + f));
        KB_TRY(select_dev(ctx, dW, nw, dth + f, dsel + (size_t)f * nw));
    }
    if (W_out) KB_CUDA(ctx, cudaMemcpyAsync(W_out, dW, sizeof(double) * nw, cudaMemcpyDeviceToHost, ctx->stream));
    if (m > 1 && kappa_out) KB_CUDA(ctx, cudaMemcpyAsync(kappa_out, dkap, sizeof(int64_t) * nw, cudaMemcpyDeviceToHost, ctx->stream));
    if (m > 1 && tau_out) KB_CUDA(ctx, cudaMemcpyAsync(tau_out, dtau, sizeof(double) * nw, cudaMemcpyDeviceToHost, ctx->stream));
    if (nfdr > 0) KB_CUDA(ctx, cudaMemcpyAsync(taus_out, dth, sizeof(double) * nfdr, cudaMemcpyDeviceToHost, ctx->stream));
    if (nfdr > 0 && selected_out)
        KB_CUDA(ctx, cudaMemcpyAsync(selected_out, dsel, sizeof(int32_t) * nw * nfdr, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_w_out) *n_w_out = nw;
    return KB_OK;
}

int check_windows(kb_ctx* ctx, const int64_t* off, int nwin, int64_t p) {
    if (!off || nwin < 1 || off[0] != 0 || off[nwin] != p)           // approx.jl:73-75
        return set_error(ctx, KB_ERR_INVALID, "window_ranges have %lld dimensions but X has %lld",
                         (long long)(off && nwin >= 1 ? off[nwin] - off[0] : 0), (long long)p);
    for (int w = 0; w < nwin; ++w)
        if (off[w + 1] <= off[w]) return set_error(ctx, KB_ERR_INVALID, "window %d is empty or out of order", w + 1);
    return KB_OK;
}
size_t window_block_doubles(const int64_t* off, int nwin) {
    size_t t = 0;
    for (int w = 0; w < nwin; ++w) t += (size_t)(off[w + 1] - off[w]) * (size_t)(off[w + 1] - off[w]);
    return t;
}
size_t window_block_offset(const int64_t* off, int w) { return window_block_doubles(off, w); }

// columns [c0, c1) of the host matrix X (n x p, ld n) -> device slab with an even leading dimension
int h2d_columns(kb_ctx* ctx, Scope& sc, const double* X, long n, long c0, long c1, double** dX, long* ldx) {
    *ldx = (n + 1) & ~1L;
    KB_TRY(sc.alloc(dX, (size_t)(*ldx) * (size_t)(c1 - c0)));
    return h2d_matrix(ctx, X + (size_t)c0 * n, n, *dX, *ldx, n, c1 - c0, ctx->stream);
}

// 1-based index array of the ABI -> 0-based vector
int from_one_based(kb_ctx* ctx, const int64_t* idx, int64_t n, int64_t p, std::vector<int>* out, const char* what) {
    out->clear();
    for (int64_t i = 0; i < n; ++i) {
        if (idx[i] < 1 || idx[i] > p) return set_error(ctx, KB_ERR_INVALID, "%s: index %lld is outside 1..%lld", what, (long long)idx[i], (long long)p);
        out->push_back((int)(idx[i] - 1));
    }
    return KB_OK;
}
// Σ[reps, reps] of Symmetric(Σ) (host)
std::vector<double> gather_sigma11(const double* Sigma, int64_t p, const std::vector<int>& reps) {
    const size_t r = reps.size();
    std::vector<double> S11(r * r);
    for (size_t j = 0; j < r; ++j)
        for (size_t i = 0; i < r; ++i) {
            const int a = std::min(reps[i], reps[j]), b = std::max(reps[i], reps[j]);
            S11[i + j * r] = Sigma[a + (size_t)b * p];
        }
    return S11;
}

}  // namespace

extern "C" {

const char* kb_version(void) { return "knockoffs_b200 0.1.0 (sm_100a)"; }

static int create_ctx(kb_ctx** out, int device, bool background);
int kb_create(kb_ctx** out, int device) { return create_ctx(out, device, false); }
int kb_create_background(kb_ctx** out, int device) { return create_ctx(out, device, true); }

static int create_ctx(kb_ctx** out, int device, bool background) {
    if (!out) return KB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return KB_ERR_CUDA;   // no CPU fallback, by design
    if (device < 0 || device >= ndev) return KB_ERR_INVALID;
    kb_ctx* ctx = new (std::nothrow) kb_ctx();
    if (!ctx) return KB_ERR_CUDA;
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    bool ok = cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok) {
        ctx->num_sms = prop.multiProcessorCount;
        ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
        ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        ctx->l2_size = (size_t)prop.l2CacheSize;
        // NOTE: the persisting-L2 carve-out is NOT enabled here: reserving it takes that capacity away from
        // normal accesses (measured: the CCD sweep went from 208 ms to 375 ms at p = 5000 with it on).
        if (getenv("KB_CCD_L2_PERSIST") && atof(getenv("KB_CCD_L2_PERSIST")) > 0.0 && ctx->l2_persist_max > 0)
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize,
                               (size_t)(atof(getenv("KB_CCD_L2_PERSIST")) * (double)ctx->l2_persist_max));
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        ok = coop != 0;
    }
    // The main stream carries the latency-bound chains (single-CTA diagonal / sequential kernels); the auxiliary stream
    // carries throughput work that runs beside them (look-ahead trailing updates, the blocked CCD's rank-256 passes).
    // Highest priority for the main stream, lowest for the auxiliary one: a freed SM slot goes to the chain first.
    int prio_least = 0, prio_greatest = 0;
    if (ok) cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    ok = ok && cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, background ? prio_least : prio_greatest) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        // KB_GREEN_SMS = SMs kept free of auxiliary-stream work (default 16; 0 = no partition)
        const char* e = getenv("KB_GREEN_SMS");
        const int reserve = e ? atoi(e) : 16;
        ctx->background = background ? 1 : 0;
        if (!(reserve > 0 && make_green_aux_stream(ctx, reserve, prio_least, background)))
            ok = cudaStreamCreateWithPriority(&ctx->stream_aux, cudaStreamNonBlocking, prio_least) == cudaSuccess;
    }
    for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i) ok = cudaEventCreateWithFlags(&ctx->ccd_ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess;
    for (int i = 0; i < 6 && ok; ++i) ok = cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        kb_destroy(ctx);
        return KB_ERR_CUDA;
    }
    *out = ctx;
    return KB_OK;
}

void kb_destroy(kb_ctx* ctx) {
    if (!ctx) return;
    {
        DeviceGuard g(ctx->device);
        if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
        for (auto& b : ctx->pool_free) cudaFree(b.ptr);
        ctx->pool_free.clear();
        release_pinned(ctx);
        if (ctx->stream_h2d) cudaStreamDestroy(ctx->stream_h2d);
        if (ctx->stream_d2h) cudaStreamDestroy(ctx->stream_d2h);
        if (ctx->stream_aux) { cudaStreamSynchronize(ctx->stream_aux); cudaStreamDestroy(ctx->stream_aux); }
        destroy_green(ctx);
        for (int i = 0; i < 2; ++i)
            if (ctx->aux_ev[i]) cudaEventDestroy(ctx->aux_ev[i]);
        for (int i = 0; i < 3; ++i)
            if (ctx->ccd_ev[i]) cudaEventDestroy(ctx->ccd_ev[i]);
        if (ctx->ev0) cudaEventDestroy(ctx->ev0);
        if (ctx->ev1) cudaEventDestroy(ctx->ev1);
        for (int i = 0; i < 6; ++i)
            if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    }
    delete ctx;
}

const char* kb_last_error(const kb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "invalid context"; }
void kb_default_solve_opts(kb_solve_opts* o) { if (o) default_solve_opts(o); }
uint64_t kb_launch_count(const kb_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* kb_stream(const kb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int32_t kb_aux_sms(const kb_ctx* ctx) { return ctx ? ctx->aux_sms : 0; }
int kb_set_sample_chunk_rows(kb_ctx* ctx, int64_t rows) {
    if (!ctx || rows < 0) return KB_ERR_INVALID;
    ctx->sample_chunk_rows = rows;
    return KB_OK;
}
int kb_set_positive_cholesky(kb_ctx* ctx, int32_t on) {
    if (!ctx) return KB_ERR_INVALID;
    ctx->positive_cholesky = on ? 1 : 0;
    return KB_OK;
}
int32_t kb_positive_modified(const kb_ctx* ctx) { return ctx ? ctx->positive_modified : 0; }
int kb_trim(kb_ctx* ctx) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& b : ctx->pool_free) cudaFree(b.ptr);
    ctx->pool_free.clear();
    ctx->pool_bytes = 0;
    release_pinned(ctx);
    return KB_OK;
}
int kb_synchronize(kb_ctx* ctx) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// ---- (1) s-solvers --------------------------------------------------------------------------------
int kb_solve_s_dev(kb_ctx* ctx, const double* dSigma, int64_t p, int32_t method, double m, const kb_solve_opts* opts,
                   const double* d_s_init, double* d_s_out, kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!dSigma || !d_s_out || p < 1 || p > 0x3fffffff) return set_error(ctx, KB_ERR_INVALID, "solve_s: bad arguments");
    DeviceGuard g(ctx->device);
    return solve_s_dev(ctx, dSigma, (long)p, (int)p, method, m, opts, d_s_init, d_s_out, info);
}

int kb_solve_s(kb_ctx* ctx, const double* Sigma, int64_t p, int32_t method, double m, const kb_solve_opts* opts,
               const double* s_init, double* s_out, kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !s_out || p < 1 || p > 0x3fffffff) return set_error(ctx, KB_ERR_INVALID, "solve_s: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dS, *ds_init = nullptr, *ds;
    KB_TRY(sc.alloc(&dS, (size_t)p * p));
    KB_TRY(sc.alloc(&ds, p));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, Sigma, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    if (s_init) {
        KB_TRY(sc.alloc(&ds_init, p));
        KB_CUDA(ctx, cudaMemcpyAsync(ds_init, s_init, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int rc = solve_s_dev(ctx, dS, (long)p, (int)p, method, m, opts, ds_init, ds, info);
    // like the reference, nothing is returned when the solver throws; s_out is only written on success
    if (rc != KB_OK) return rc;
    KB_CUDA(ctx, cudaMemcpyAsync(s_out, ds, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_eigmin(kb_ctx* ctx, const double* Sigma, int64_t p, double* lambda_min_out) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !lambda_min_out || p < 1) return set_error(ctx, KB_ERR_INVALID, "eigmin: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double* dS;
    KB_TRY(sc.alloc(&dS, (size_t)p * p));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, Sigma, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    return eigmin_dev(ctx, dS, (long)p, (int)p, lambda_min_out);
}

// ---- (1b) group s-solvers ---------------------------------------------------------------------------
void kb_default_group_opts(kb_group_opts* o) { if (o) default_group_opts(o); }

int kb_solve_s_group(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, int32_t method, double m,
                     const kb_group_opts* opts, const double* V, int64_t nv, double* S_out, double* obj_out,
                     kb_group_info* info) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !groups || !S_out || p < 1 || p > 46000 || nv < 0 || nv > 0x3fffffff)
        return set_error(ctx, KB_ERR_INVALID, "solve_s_group: bad arguments");
    DeviceGuard g(ctx->device);
    return solve_s_group_host(ctx, Sigma, (int)p, groups, method, m, opts, V, (int)nv, nullptr, nullptr, S_out, obj_out, info);
}

int kb_solve_s_group_sharded(kb_ctx* ctx, const double* Sigma_block, int64_t p, const int64_t* groups, int32_t method,
                             double m, const kb_group_opts* opts, const double* V, int64_t nv, kb_reduce_fn reduce,
                             void* user, double* S_out, double* obj_local_out, kb_group_info* info) {
    if (!ctx) return bad_ctx();
    if (!Sigma_block || !groups || !S_out || !reduce || p < 1 || p > 46000 || nv < 0 || nv > 0x3fffffff)
        return set_error(ctx, KB_ERR_INVALID, "solve_s_group_sharded: bad arguments");
    DeviceGuard g(ctx->device);
    return solve_s_group_host(ctx, Sigma_block, (int)p, groups, method, m, opts, V, (int)nv, reduce, user, S_out,
                              obj_local_out, info);
}

int kb_modelx_gaussian_group_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* groups,
                                       const double* mu, const double* Sigma, int32_t method, int32_t m,
                                       const kb_group_opts* opts, const double* V, int64_t nv, const double* Z,
                                       uint64_t seed, int64_t row_offset, double* Xko_out, double* S_out,
                                       double* obj_out, kb_group_info* info) {
    if (!ctx) return bad_ctx();
    if (!X || !mu || !Sigma || !groups || !Xko_out || !S_out || p < 1 || n < 0)
        return set_error(ctx, KB_ERR_INVALID, "modelX group: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    // (S, γs, obj) = solve_s_group(Symmetric(Σ), groups, method; m, kwargs...)      (group.jl:123)
    KB_TRY(solve_s_group_host(ctx, Sigma, (int)p, groups, method, (double)m, opts, V, (int)nv, nullptr, nullptr, S_out,
                              obj_out, info));
    // X̃ = condition(X, μ, Σ, S; m)                                                  (group.jl:125)
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dSig, *dS, *dmu, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&dS, pp));
    KB_TRY(sc.alloc(&dmu, p));
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_CUDA(ctx, cudaMemcpyAsync(dSig, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, S_out, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(cond_factor_dev(ctx, dSig, (long)p, (int)p, dS, (long)p, 0, m, dSiS, dLb, (long)p));
    return sample_host_pipeline(ctx, X, (long)n, (int)p, dmu, dSiS, dLb, (long)p, m, Z, seed, row_offset, Xko_out);
}

// ---- (2) Gram -------------------------------------------------------------------------------------
int kb_gram_partial_dev(kb_ctx* ctx, const double* dX, int64_t n_local, int64_t ldx, int64_t p, double* dG,
                        double* dcolsum) {
    if (!ctx) return bad_ctx();
    if (!dX || !dG || p < 1 || n_local < 0 || ldx < n_local) return set_error(ctx, KB_ERR_INVALID, "gram_partial: bad arguments");
    DeviceGuard g(ctx->device);
    KB_TRY(fill_zero(ctx, dG, (size_t)p * p));
    if (dcolsum) KB_TRY(fill_zero(ctx, dcolsum, p));
    KB_TRY(gram_partial_dev(ctx, dX, (long)n_local, (long)ldx, (int)p, dG, (long)p, dcolsum, 0));
    KB_TRY(mirror_lower_to_upper(ctx, dG, (long)p, (int)p));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_gram_finish_dev(kb_ctx* ctx, double* dG, double* dcolsum, int64_t n_total, int64_t p, int32_t center,
                       double denom) {
    if (!ctx) return bad_ctx();
    if (!dG || p < 1 || n_total < 1) return set_error(ctx, KB_ERR_INVALID, "gram_finish: bad arguments");
    DeviceGuard g(ctx->device);
    KB_TRY(gram_finish_dev(ctx, dG, (long)p, dcolsum, (long)n_total, (int)p, center, denom));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_gram(kb_ctx* ctx, const double* X, int64_t n, int64_t p, int32_t center, double denom, double* G_out,
            double* colmean_out) {
    if (!ctx) return bad_ctx();
    if (!X || !G_out || p < 1 || n < 1) return set_error(ctx, KB_ERR_INVALID, "gram: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const long ld = (p + 1) & ~1L;
    double *dG, *dcs;
    KB_TRY(sc.alloc(&dG, (size_t)ld * p));
    KB_TRY(sc.alloc(&dcs, p));
    KB_TRY(fill_zero(ctx, dG, (size_t)ld * p));
    KB_TRY(fill_zero(ctx, dcs, p));
    KB_TRY(gram_accumulate_host_rows(ctx, X, (long)n, (long)n, (int)p, dG, ld, dcs));
    KB_TRY(gram_finish_dev(ctx, dG, ld, dcs, (long)n, (int)p, center, denom));
    KB_TRY(d2h_matrix(ctx, dG, ld, G_out, p, p, p, ctx->stream));
    if (colmean_out) KB_CUDA(ctx, cudaMemcpyAsync(colmean_out, dcs, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_h2d));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// Σ̂ = cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X), μ̂ = column means  (modelX.jl:43-50).  Two passes
// over the host matrix: (1) column sums + raw Gram, (2) Gram of the squared centred entries.
int kb_cov_shrink_lw(kb_ctx* ctx, const double* X, int64_t n, int64_t p, double* Sigma_out, double* colmean_out,
                     double* shrinkage_out) {
    if (!ctx) return bad_ctx();
    if (!X || !Sigma_out || p < 1 || n < 2) return set_error(ctx, KB_ERR_INVALID, "cov_shrink_lw: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    StreamDrain drain(ctx);
    const long ld = (p + 1) & ~1L;
    long chunk = 32768;
    if (chunk > n) chunk = (n + 1) & ~1L;
    double *dG, *dG2, *dcs, *dXc[2], *dY;
    KB_TRY(sc.alloc(&dG, (size_t)ld * p));
    KB_TRY(sc.alloc(&dG2, (size_t)ld * p));
    KB_TRY(sc.alloc(&dcs, p));
    KB_TRY(sc.alloc(&dY, (size_t)chunk * p));
    for (int b = 0; b < 2; ++b) KB_TRY(sc.alloc(&dXc[b], (size_t)chunk * p));
    KB_TRY(fill_zero(ctx, dG, (size_t)ld * p));
    KB_TRY(fill_zero(ctx, dG2, (size_t)ld * p));
    KB_TRY(fill_zero(ctx, dcs, p));
    cudaEvent_t* ev_h2d = ctx->pipe_ev;
    cudaEvent_t* ev_comp = ctx->pipe_ev + 2;
    for (int pass = 0; pass < 2; ++pass) {
        long i = 0;
        for (long r0 = 0; r0 < n; r0 += chunk, ++i) {
            const int b = (int)(i & 1);
            const long rows = std::min(chunk, (long)n - r0);
            if (i >= 2) KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ev_comp[b], 0));
            KB_TRY(h2d_matrix(ctx, X + r0, n, dXc[b], chunk, rows, p, ctx->stream_h2d));
            KB_CUDA(ctx, cudaEventRecord(ev_h2d[b], ctx->stream_h2d));
            KB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[b], 0));
            if (pass == 0) KB_TRY(gram_partial_dev(ctx, dXc[b], rows, chunk, (int)p, dG, ld, dcs, 1));
            else KB_TRY(gram_sq_partial_dev(ctx, dXc[b], rows, chunk, (int)p, dcs, dY, chunk, dG2, ld, 1));
            KB_CUDA(ctx, cudaEventRecord(ev_comp[b], ctx->stream));
        }
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream_h2d));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pass == 0) KB_TRY(gram_finish_dev(ctx, dG, ld, dcs, (long)n, (int)p, 1, (double)n));   // S = Xc'Xc/n, dcs -> means
    }
    double lam = 0.0;
    KB_TRY(lw_finish_dev(ctx, dG, ld, dG2, ld, (int)p, (long)n, &lam));
    KB_TRY(d2h_matrix(ctx, dG, ld, Sigma_out, p, p, p, ctx->stream));
    if (colmean_out) KB_CUDA(ctx, cudaMemcpyAsync(colmean_out, dcs, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (shrinkage_out) *shrinkage_out = lam;
    return KB_OK;
}

// ---- (3) conditional factor -----------------------------------------------------------------------
int kb_cond_factor_dev(kb_ctx* ctx, const double* dSigma, int64_t p, const double* dS, int32_t s_is_diag, int32_t m,
                       double* dSigmaInvS, double* dLblocks) {
    if (!ctx) return bad_ctx();
    if (!dSigma || !dS || !dSigmaInvS || !dLblocks || p < 1) return set_error(ctx, KB_ERR_INVALID, "cond_factor: bad arguments");
    DeviceGuard g(ctx->device);
    KB_TRY(cond_factor_dev(ctx, dSigma, (long)p, (int)p, dS, (long)p, s_is_diag, m, dSigmaInvS, dLblocks, (long)p));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_cond_factor_block_dev(kb_ctx* ctx, const double* dSigma, int64_t p, const double* ds, int32_t k, int32_t m,
                             double* dSigmaInvS, double* dLkk, double* dMk) {
    if (!ctx) return bad_ctx();
    if (!dSigma || !ds || !dLkk || p < 1) return set_error(ctx, KB_ERR_INVALID, "cond_factor_block: bad arguments");
    DeviceGuard g(ctx->device);
    KB_TRY(cond_factor_block_dev(ctx, dSigma, (long)p, (int)p, ds, k, m, dSigmaInvS, dLkk, dMk, (long)p));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_cond_factor(kb_ctx* ctx, const double* Sigma, int64_t p, const double* S, int32_t s_is_diag, int32_t m,
                   double* SigmaInvS_out, double* Lblocks_out) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !S || !SigmaInvS_out || !Lblocks_out || p < 1) return set_error(ctx, KB_ERR_INVALID, "cond_factor: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dSig, *dS, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&dS, s_is_diag ? (size_t)p : pp));
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_CUDA(ctx, cudaMemcpyAsync(dSig, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, S, sizeof(double) * (s_is_diag ? (size_t)p : pp), cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(cond_factor_dev(ctx, dSig, (long)p, (int)p, dS, (long)p, s_is_diag, m, dSiS, dLb, (long)p));
    KB_CUDA(ctx, cudaMemcpyAsync(SigmaInvS_out, dSiS, sizeof(double) * pp, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(Lblocks_out, dLb, sizeof(double) * pp * nblk, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// ---- (4) sampling ---------------------------------------------------------------------------------
int kb_sample_dev(kb_ctx* ctx, const double* dX, int64_t n, int64_t ldx, int64_t p, const double* dmu,
                  const double* dSigmaInvS, const double* dLblocks, int32_t m, const double* dZ, int64_t ldz,
                  uint64_t seed, int64_t row_offset, double* dXko, int64_t ldxko) {
    if (!ctx) return bad_ctx();
    if (!dX || !dmu || !dSigmaInvS || !dLblocks || !dXko || p < 1 || n < 0 || ldx < n || ldxko < n || (dZ && ldz < n))
        return set_error(ctx, KB_ERR_INVALID, "sample: bad arguments");
    DeviceGuard g(ctx->device);
    return sample_dev(ctx, dX, (long)n, (long)ldx, (int)p, dmu, dSigmaInvS, dLblocks, (long)p, m, dZ, (long)ldz, seed,
                      row_offset, dXko, (long)ldxko);
}

int kb_sample(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu, const double* SigmaInvS,
              const double* Lblocks, int32_t m, const double* Z, uint64_t seed, int64_t row_offset, double* Xko_out) {
    if (!ctx) return bad_ctx();
    if (!X || !mu || !SigmaInvS || !Lblocks || !Xko_out || p < 1 || n < 0) return set_error(ctx, KB_ERR_INVALID, "sample: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dmu, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dmu, p));
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dSiS, SigmaInvS, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dLb, Lblocks, sizeof(double) * pp * nblk, cudaMemcpyHostToDevice, ctx->stream));
    return sample_host_pipeline(ctx, X, (long)n, (int)p, dmu, dSiS, dLb, (long)p, m, Z, seed, row_offset, Xko_out);
}

int kb_condition(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu, const double* Sigma,
                 const double* S, int32_t s_is_diag, int32_t m, const double* Z, uint64_t seed, int64_t row_offset,
                 double* Xko_out) {
    if (!ctx) return bad_ctx();
    if (!X || !mu || !Sigma || !S || !Xko_out || p < 1 || n < 0) return set_error(ctx, KB_ERR_INVALID, "condition: bad arguments");
    DeviceGuard g(ctx->device);
    return condition_host(ctx, X, (long)n, (long)n, (int)p, mu, Sigma, S, s_is_diag, m, Z, seed, row_offset, Xko_out);
}

// ---- one-shot ---------------------------------------------------------------------------------------
int kb_modelx_gaussian_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const double* mu,
                                 const double* Sigma, int32_t method, int32_t m, const kb_solve_opts* opts,
                                 const double* s_init, const double* Z, uint64_t seed, int64_t row_offset,
                                 double* Xko_out, double* s_out, kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!X || !mu || !Sigma || !Xko_out || !s_out || p < 1 || n < 0) return set_error(ctx, KB_ERR_INVALID, "modelX: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dSig, *ds, *ds_init = nullptr, *dmu, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&ds, p));
    KB_TRY(sc.alloc(&dmu, p));
    KB_CUDA(ctx, cudaMemcpyAsync(dSig, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    if (s_init) {
        KB_TRY(sc.alloc(&ds_init, p));
        KB_CUDA(ctx, cudaMemcpyAsync(ds_init, s_init, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    }
    // s = solve_s(Symmetric(Σ), method; m, kwargs...)                    (modelX.jl:62)
    KB_TRY(solve_s_dev(ctx, dSig, (long)p, (int)p, method, (double)m, opts, ds_init, ds, info));
    KB_CUDA(ctx, cudaMemcpyAsync(s_out, ds, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    // X̃ = condition(X, μ, Symmetric(Σ), Diagonal(s); m)                 (modelX.jl:64)
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_TRY(cond_factor_dev(ctx, dSig, (long)p, (int)p, ds, (long)p, 1, m, dSiS, dLb, (long)p));
    return sample_host_pipeline(ctx, X, (long)n, (int)p, dmu, dSiS, dLb, (long)p, m, Z, seed, row_offset, Xko_out);
}


// ---- (5) feature statistics and the knockoff filter ----------------------------------
int kb_marginal_stats_dev(kb_ctx* ctx, const double* dy, const double* dX, int64_t ldx, const double* dXko, int64_t ldxko,
                          int64_t n, int64_t p, int32_t m, double* dT_out) {
    if (!ctx) return bad_ctx();
    if (!dy || !dX || !dXko || !dT_out || p < 1 || n < 2 || m < 1 || ldx < n || ldxko < n)
        return set_error(ctx, KB_ERR_INVALID, "marginal_stats: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dys, *dsum, *dacc;
    KB_TRY(sc.alloc(&dys, (size_t)n));
    KB_TRY(sc.alloc(&dsum, 4));
    KB_TRY(sc.alloc(&dacc, (size_t)STATS_ACC * (m + 1) * p));
    KB_TRY(zscore_vec_dev(ctx, dy, (long)n, dys, dsum));
    KB_TRY(marginal_accumulate_dev(ctx, dX, (long)n, (long)ldx, (long)p, dys, dacc, 1));
    KB_TRY(marginal_accumulate_dev(ctx, dXko, (long)n, (long)ldxko, (long)p * m, dys, dacc + (size_t)STATS_ACC * p, 1));
    KB_TRY(marginal_finish_dev(ctx, dacc, (long)(m + 1) * p, (long)n, dsum, dT_out));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_marginal_stats(kb_ctx* ctx, const double* y, const double* X, const double* Xko, int64_t n, int64_t p, int32_t m,
                      double* T_out) {
    if (!ctx) return bad_ctx();
    if (!y || !X || !Xko || !T_out || p < 1 || n < 2 || m < 1) return set_error(ctx, KB_ERR_INVALID, "marginal_stats: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const long nc = (long)(m + 1) * p;
    double *dy, *dys, *dsum, *dacc, *dT;
    KB_TRY(sc.alloc(&dy, (size_t)n));
    KB_TRY(sc.alloc(&dys, (size_t)n));
    KB_TRY(sc.alloc(&dsum, 4));
    KB_TRY(sc.alloc(&dacc, (size_t)STATS_ACC * nc));
    KB_TRY(sc.alloc(&dT, (size_t)nc));
    KB_CUDA(ctx, cudaMemcpyAsync(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(zscore_vec_dev(ctx, dy, (long)n, dys, dsum));
    KB_TRY(host_marginal_accumulate(ctx, X, (long)n, (long)p, dys, dacc));
    KB_TRY(host_marginal_accumulate(ctx, Xko, (long)n, (long)p * m, dys, dacc + (size_t)STATS_ACC * p));
    KB_TRY(marginal_finish_dev(ctx, dacc, nc, (long)n, dsum, dT));
    KB_CUDA(ctx, cudaMemcpyAsync(T_out, dT, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_mk_statistics(kb_ctx* ctx, const double* T0, const double* Tk, int64_t p, int32_t m, int64_t* kappa_out,
                     double* tau_out, double* W_out) {
    if (!ctx) return bad_ctx();
    if (!T0 || !Tk || !W_out || p < 0 || m < 1 || (m > 1 && (!kappa_out || !tau_out)))
        return set_error(ctx, KB_ERR_INVALID, "MK_statistics: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t pp = (size_t)std::max<int64_t>(1, p);
    double *dT0, *dTk, *dtau, *dW;
    long long* dkap;
    KB_TRY(sc.alloc(&dT0, pp));
    KB_TRY(sc.alloc(&dTk, pp * m));
    KB_TRY(sc.alloc(&dtau, pp));
    KB_TRY(sc.alloc(&dW, pp));
    KB_TRY(sc.alloc(&dkap, pp));
    KB_CUDA(ctx, cudaMemcpyAsync(dT0, T0, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dTk, Tk, sizeof(double) * p * m, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(mk_statistics_dev(ctx, dT0, dTk, (long)p, m, dkap, dtau, dW));
    KB_CUDA(ctx, cudaMemcpyAsync(W_out, dW, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    if (m > 1) {
        KB_CUDA(ctx, cudaMemcpyAsync(kappa_out, dkap, sizeof(int64_t) * p, cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaMemcpyAsync(tau_out, dtau, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_threshold(kb_ctx* ctx, const double* w, int64_t p, double q, int32_t knockoff_plus, int64_t rej_bounds,
                 double* tau_out) {
    if (!ctx) return bad_ctx();
    if (!w || !tau_out || p < 0) return set_error(ctx, KB_ERR_INVALID, "threshold: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dw, *dout;
    KB_TRY(sc.alloc(&dw, (size_t)std::max<int64_t>(1, p)));
    KB_TRY(sc.alloc(&dout, 4));
    KB_CUDA(ctx, cudaMemcpyAsync(dw, w, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(threshold_dev(ctx, dw, (long)p, q, knockoff_plus ? 1 : 0, (long)rej_bounds, dout));
    KB_CUDA(ctx, cudaMemcpyAsync(tau_out, dout, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_mk_threshold(kb_ctx* ctx, const double* tau, const int64_t* kappa, int64_t p, int32_t m, double q,
                    int64_t rej_bounds, double* tau_hat_out) {
    if (!ctx) return bad_ctx();
    if (!tau || !kappa || !tau_hat_out || p < 0) return set_error(ctx, KB_ERR_INVALID, "mk_threshold: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dtau, *dout;
    long long* dkap;
    KB_TRY(sc.alloc(&dtau, (size_t)std::max<int64_t>(1, p)));
    KB_TRY(sc.alloc(&dkap, (size_t)std::max<int64_t>(1, p)));
    KB_TRY(sc.alloc(&dout, 4));
    KB_CUDA(ctx, cudaMemcpyAsync(dtau, tau, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dkap, kappa, sizeof(int64_t) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(mk_threshold_dev(ctx, dtau, dkap, (long)p, m, q, (long)rej_bounds, dout));
    KB_CUDA(ctx, cudaMemcpyAsync(tau_hat_out, dout, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_marginal_filter(kb_ctx* ctx, const double* y, const double* X, const double* Xko, int64_t n, int64_t p, int32_t m,
                       const int64_t* groups, const double* fdrs, int32_t nfdr, int32_t knockoff_plus, double* T_out,
                       double* W_out, int64_t* kappa_out, double* tau_out, double* taus_out, int32_t* selected_out,
                       int64_t* n_w_out) {
    if (!ctx) return bad_ctx();
    if (!y || !X || !Xko || p < 1 || n < 2 || m < 1) return set_error(ctx, KB_ERR_INVALID, "marginal_filter: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const long nc = (long)(m + 1) * p;
    double *dy, *dys, *dsum, *dacc, *dT;
    KB_TRY(sc.alloc(&dy, (size_t)n));
    KB_TRY(sc.alloc(&dys, (size_t)n));
    KB_TRY(sc.alloc(&dsum, 4));
    KB_TRY(sc.alloc(&dacc, (size_t)STATS_ACC * nc));
    KB_TRY(sc.alloc(&dT, (size_t)nc));
    KB_CUDA(ctx, cudaMemcpyAsync(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(zscore_vec_dev(ctx, dy, (long)n, dys, dsum));
    KB_TRY(host_marginal_accumulate(ctx, X, (long)n, (long)p, dys, dacc));
    KB_TRY(host_marginal_accumulate(ctx, Xko, (long)n, (long)p * m, dys, dacc + (size_t)STATS_ACC * p));
    KB_TRY(marginal_finish_dev(ctx, dacc, nc, (long)n, dsum, dT));
    if (T_out) KB_CUDA(ctx, cudaMemcpyAsync(T_out, dT, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
    return knockoff_filter_from_T(ctx, dT, (long)p, m, groups, fdrs, nfdr, knockoff_plus, 10000, W_out, kappa_out, tau_out,
                                  taus_out, selected_out, n_w_out);
}

int kb_fit_marginal(kb_ctx* ctx, const double* y, const double* X, int64_t n, int64_t p, const double* mu,
                    const double* Sigma, int32_t method, int32_t m, const kb_solve_opts* opts, const double* fdrs,
                    int32_t nfdr, int32_t knockoff_plus, const double* Z, uint64_t seed, double* Xko_out, double* s_out,
                    double* T_out, double* W_out, int64_t* kappa_out, double* tau_out, double* taus_out,
                    int32_t* selected_out, kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!y || !X || !mu || !Sigma || p < 1 || n < 2) return set_error(ctx, KB_ERR_INVALID, "fit_marginal: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    const long nc = (long)(m + 1) * p;
    double *dSig, *ds, *dmu, *dSiS, *dLb, *dy, *dys, *dsum, *dacc, *dT;
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&ds, p));
    KB_TRY(sc.alloc(&dmu, p));
    KB_TRY(sc.alloc(&dy, (size_t)n));
    KB_TRY(sc.alloc(&dys, (size_t)n));
    KB_TRY(sc.alloc(&dsum, 4));
    KB_TRY(sc.alloc(&dacc, (size_t)STATS_ACC * nc));
    KB_TRY(sc.alloc(&dT, (size_t)nc));
    KB_CUDA(ctx, cudaMemcpyAsync(dSig, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    // ko = modelX_gaussian_knockoffs(X, method, μ, Σ; m)                  (fit_lasso.jl:199)
    KB_TRY(solve_s_dev(ctx, dSig, (long)p, (int)p, method, (double)m, opts, nullptr, ds, info));
    if (s_out) KB_CUDA(ctx, cudaMemcpyAsync(s_out, ds, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_TRY(cond_factor_dev(ctx, dSig, (long)p, (int)p, ds, (long)p, 1, m, dSiS, dLb, (long)p));
    // feature importance while the chunks of X and X̃ are on the device   (fit_lasso.jl:214-222)
    KB_TRY(zscore_vec_dev(ctx, dy, (long)n, dys, dsum));
    StatsHook hook{dys, dacc, dacc + (size_t)STATS_ACC * p};
    KB_TRY(sample_host_pipeline(ctx, X, (long)n, (int)p, dmu, dSiS, dLb, (long)p, m, Z, seed, 0, Xko_out, &hook));
    KB_TRY(marginal_finish_dev(ctx, dacc, nc, (long)n, dsum, dT));
    if (T_out) KB_CUDA(ctx, cudaMemcpyAsync(T_out, dT, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
    return knockoff_filter_from_T(ctx, dT, (long)p, m, nullptr, fdrs, nfdr, knockoff_plus, 10000, W_out, kappa_out, tau_out,
                                  taus_out, selected_out, nullptr);
}


// ---- (6) approx_modelX_gaussian_knockoffs ---------------------------------------------
int kb_feasible_gamma(kb_ctx* ctx, const double* Sigma, int64_t p, const double* s, double m, double* gamma_out) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !s || !gamma_out || p < 1 || !(m >= 1.0)) return set_error(ctx, KB_ERR_INVALID, "feasible_gamma: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dS, *ds;
    KB_TRY(sc.alloc(&dS, (size_t)p * p));
    KB_TRY(sc.alloc(&ds, (size_t)p));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, Sigma, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(ds, s, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    return feasible_gamma_dev(ctx, dS, (long)p, (int)p, ds, m, gamma_out);
}

int kb_approx_solve_windows(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* window_offsets,
                            int32_t nwin, int32_t w0, int32_t w1, int32_t method, int32_t m, const kb_solve_opts* opts,
                            double* Sigma_blocks_out, double* s_out, double* mu_out, double* gamma_min_out,
                            kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!X || !Sigma_blocks_out || !s_out || !mu_out || !gamma_min_out || p < 1 || n < 2)
        return set_error(ctx, KB_ERR_INVALID, "approx_solve_windows: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    KB_TRY(check_windows(ctx, window_offsets, nwin, p));
    if (w0 < 0 || w1 > nwin || w0 > w1) return set_error(ctx, KB_ERR_INVALID, "approx_solve_windows: bad window range");
    *gamma_min_out = 1.0;
    if (w0 == w1) return KB_OK;
    Scope sc(ctx);
    const long c0 = (long)window_offsets[w0], c1 = (long)window_offsets[w1];
    double *dX, *dblk, *ds, *dmu;
    long ldx;
    KB_TRY(h2d_columns(ctx, sc, X, (long)n, c0, c1, &dX, &ldx));
    KB_TRY(sc.alloc(&dblk, window_block_doubles(window_offsets, nwin)));
    KB_TRY(sc.alloc(&ds, (size_t)p));
    KB_TRY(sc.alloc(&dmu, (size_t)p));
    KB_TRY(approx_solve_windows(ctx, dX, (long)n, ldx, (long)p, window_offsets, nwin, w0, w1, method, (double)m, opts, dblk, ds,
                                dmu, gamma_min_out, info));
    const size_t b0 = window_block_offset(window_offsets, w0), b1 = window_block_offset(window_offsets, w1);
    KB_CUDA(ctx, cudaMemcpyAsync(Sigma_blocks_out + b0, dblk + b0, sizeof(double) * (b1 - b0), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(s_out + c0, ds + c0, sizeof(double) * (c1 - c0), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(mu_out + c0, dmu + c0, sizeof(double) * (c1 - c0), cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_approx_condition_windows(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* window_offsets,
                                int32_t nwin, int32_t w0, int32_t w1, int32_t m, const double* Sigma_blocks,
                                const double* s, const double* mu, const double* Z, uint64_t seed, int64_t row_offset,
                                double* Xko_out) {
    if (!ctx) return bad_ctx();
    if (!X || !Sigma_blocks || !s || !mu || !Xko_out || p < 1 || n < 1)
        return set_error(ctx, KB_ERR_INVALID, "approx_condition_windows: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    KB_TRY(check_windows(ctx, window_offsets, nwin, p));
    if (w0 < 0 || w1 > nwin || w0 > w1) return set_error(ctx, KB_ERR_INVALID, "approx_condition_windows: bad window range");
    if (w0 == w1) return KB_OK;
    Scope sc(ctx);
    const long c0 = (long)window_offsets[w0], c1 = (long)window_offsets[w1];
    const size_t b0 = window_block_offset(window_offsets, w0), b1 = window_block_offset(window_offsets, w1);
    double *dX, *dblk, *ds, *dmu;
    long ldx;
    KB_TRY(h2d_columns(ctx, sc, X, (long)n, c0, c1, &dX, &ldx));
    KB_TRY(sc.alloc(&dblk, window_block_doubles(window_offsets, nwin)));
    KB_TRY(sc.alloc(&ds, (size_t)p));
    KB_TRY(sc.alloc(&dmu, (size_t)p));
    KB_CUDA(ctx, cudaMemcpyAsync(dblk + b0, Sigma_blocks + b0, sizeof(double) * (b1 - b0), cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(ds + c0, s + c0, sizeof(double) * (c1 - c0), cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu + c0, mu + c0, sizeof(double) * (c1 - c0), cudaMemcpyHostToDevice, ctx->stream));
    return approx_condition_windows(ctx, dX, (long)n, ldx, (long)p, window_offsets, nwin, w0, w1, m, dblk, ds, dmu, Z, seed,
                                    row_offset, Xko_out);
}

int kb_approx_modelx_gaussian_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* window_offsets,
                                        int32_t nwin, int32_t method, int32_t m, const kb_solve_opts* opts, const double* Z,
                                        uint64_t seed, double* Xko_out, double* s_out, double* mu_out,
                                        double* Sigma_blocks_out, double* gamma_out, kb_solve_info* info) {
    if (!ctx) return bad_ctx();
    if (!X || !Xko_out || !s_out || p < 1 || n < 2) return set_error(ctx, KB_ERR_INVALID, "approx modelX: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    KB_TRY(check_windows(ctx, window_offsets, nwin, p));
    Scope sc(ctx);
    const size_t nblk = window_block_doubles(window_offsets, nwin);
    double *dX, *dblk, *ds, *dmu;
    long ldx;
    KB_TRY(h2d_columns(ctx, sc, X, (long)n, 0, (long)p, &dX, &ldx));
    KB_TRY(sc.alloc(&dblk, nblk));
    KB_TRY(sc.alloc(&ds, (size_t)p));
    KB_TRY(sc.alloc(&dmu, (size_t)p));
    double gamma = 1.0;
    KB_TRY(approx_solve_windows(ctx, dX, (long)n, ldx, (long)p, window_offsets, nwin, 0, nwin, method, (double)m, opts, dblk, ds,
                                dmu, &gamma, info));
    KB_TRY(scale_vec_dev(ctx, ds, (long)p, gamma));                                          // s .*= γ  (approx.jl:98)
    KB_CUDA(ctx, cudaMemcpyAsync(s_out, ds, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    if (mu_out) KB_CUDA(ctx, cudaMemcpyAsync(mu_out, dmu, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    if (Sigma_blocks_out)
        KB_CUDA(ctx, cudaMemcpyAsync(Sigma_blocks_out, dblk, sizeof(double) * nblk, cudaMemcpyDeviceToHost, ctx->stream));
    if (gamma_out) *gamma_out = gamma;
    return approx_condition_windows(ctx, dX, (long)n, ldx, (long)p, window_offsets, nwin, 0, nwin, m, dblk, ds, dmu, Z, seed, 0,
                                    Xko_out);
}


// ---- (7) representative group knockoffs ------------------------------------------------
int kb_choose_group_reps(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, double threshold,
                         const int64_t* prioritize_idx, int64_t n_prioritize, int64_t* group_reps_out,
                         int64_t* n_reps_out) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !groups || !group_reps_out || !n_reps_out || p < 1 || p > 46000 || n_prioritize < 0)
        return set_error(ctx, KB_ERR_INVALID, "choose_group_reps: bad arguments");
    DeviceGuard g(ctx->device);
    std::vector<int> prior, reps;
    if (prioritize_idx) KB_TRY(from_one_based(ctx, prioritize_idx, n_prioritize, p, &prior, "prioritize_idx"));
    KB_TRY(choose_group_reps_host(ctx, Sigma, (long)p, (int)p, groups, threshold, prioritize_idx != nullptr, prior, &reps));
    for (size_t i = 0; i < reps.size(); ++i) group_reps_out[i] = reps[i] + 1;
    *n_reps_out = (int64_t)reps.size();
    return KB_OK;
}

int kb_cond_indep_corr(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups, const int64_t* group_reps,
                       int64_t n_reps, double* Sigma_new_out) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !groups || !group_reps || !Sigma_new_out || p < 1 || n_reps < 1 || n_reps > p)
        return set_error(ctx, KB_ERR_INVALID, "cond_indep_corr: bad arguments");
    DeviceGuard g(ctx->device);
    std::vector<int> reps;
    KB_TRY(from_one_based(ctx, group_reps, n_reps, p, &reps, "group_reps"));
    Scope sc(ctx);
    double* dOut;
    KB_TRY(sc.alloc(&dOut, (size_t)p * p));
    KB_TRY(cond_indep_corr_dev(ctx, Sigma, (long)p, (int)p, groups, reps, dOut));
    KB_CUDA(ctx, cudaMemcpyAsync(Sigma_new_out, dOut, sizeof(double) * p * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_solve_s_graphical_group(kb_ctx* ctx, const double* Sigma, int64_t p, const int64_t* groups,
                               const int64_t* group_reps, int64_t n_reps, int32_t method, double m,
                               const kb_group_opts* opts, const double* V, int64_t nv, double* S_out, double* D_out,
                               double* obj_out, kb_group_info* info) {
    if (!ctx) return bad_ctx();
    if (!Sigma || !groups || !group_reps || !S_out || !D_out || p < 1 || p > 46000 || n_reps < 1 || n_reps > p || nv < 0)
        return set_error(ctx, KB_ERR_INVALID, "solve_s_graphical_group: bad arguments");
    DeviceGuard g(ctx->device);
    std::vector<int> reps;
    KB_TRY(from_one_based(ctx, group_reps, n_reps, p, &reps, "group_reps"));
    Scope sc(ctx);
    double *dRaw, *dSig, *dD;
    KB_TRY(sc.alloc(&dRaw, (size_t)p * p));
    KB_TRY(sc.alloc(&dSig, (size_t)p * p));
    KB_TRY(sc.alloc(&dD, (size_t)p * p));
    KB_CUDA(ctx, cudaMemcpyAsync(dRaw, Sigma, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(sym_upper_to_full_dev(ctx, dRaw, (long)p, (int)p, dSig, (long)p));
    std::vector<double> S11 = gather_sigma11(Sigma, p, reps);
    double obj = 0.0;
    KB_TRY(solve_s_graphical_group_dev(ctx, dSig, (int)p, S11.data(), groups, reps, method, m, opts, V, (int)nv, S_out, dD, &obj,
                                       info));
    KB_CUDA(ctx, cudaMemcpyAsync(D_out, dD, sizeof(double) * p * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (obj_out) *obj_out = obj;
    return KB_OK;
}

int kb_modelx_gaussian_rep_group_knockoffs(kb_ctx* ctx, const double* X, int64_t n, int64_t p, const int64_t* groups,
                                           const double* mu, const double* Sigma, int32_t method, int32_t m,
                                           double rep_threshold, int32_t enforce_cond_indep, const kb_group_opts* opts,
                                           const int64_t* group_reps_in, int64_t n_reps_in, const double* V, int64_t nv,
                                           const double* Z, uint64_t seed, double* Xko_out, int64_t* group_reps_out,
                                           int64_t* n_reps_out, double* S11_out, double* D_out, double* obj_out,
                                           kb_group_info* info) {
    if (!ctx) return bad_ctx();
    if (!X || !mu || !Sigma || !groups || !Xko_out || !group_reps_out || !n_reps_out || !S11_out || p < 1 || p > 46000 || n < 0 ||
        nv < 0)
        return set_error(ctx, KB_ERR_INVALID, "modelX rep group: bad arguments");
    if (m < 1) return set_error(ctx, KB_ERR_INVALID, "m should be 1 or larger but was %d.", m);
    DeviceGuard g(ctx->device);
    // group_reps = choose_group_reps(Symmetric(Σ), groups, threshold=rep_threshold)                (group.jl:186)
    std::vector<int> reps;
    if (group_reps_in) KB_TRY(from_one_based(ctx, group_reps_in, n_reps_in, p, &reps, "group_reps"));
    else KB_TRY(choose_group_reps_host(ctx, Sigma, (long)p, (int)p, groups, rep_threshold, false, {}, &reps));
    for (size_t i = 0; i < reps.size(); ++i) group_reps_out[i] = reps[i] + 1;
    *n_reps_out = (int64_t)reps.size();
    Scope sc(ctx);
    const size_t pp = (size_t)p * p;
    const size_t nblk = (size_t)(2 * m - 1);
    double *dRaw, *dSig, *dD, *dmu, *dSiS, *dLb;
    KB_TRY(sc.alloc(&dRaw, pp));
    KB_TRY(sc.alloc(&dSig, pp));
    KB_TRY(sc.alloc(&dD, pp));
    KB_TRY(sc.alloc(&dmu, (size_t)p));
    KB_CUDA(ctx, cudaMemcpyAsync(dRaw, Sigma, sizeof(double) * pp, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dmu, mu, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    // sigma = enforce_cond_indep ? cond_indep_corr(Σ, groups, group_reps) : Σ                       (group.jl:189)
    if (enforce_cond_indep) KB_TRY(cond_indep_corr_dev(ctx, Sigma, (long)p, (int)p, groups, reps, dSig));
    else KB_TRY(sym_upper_to_full_dev(ctx, dRaw, (long)p, (int)p, dSig, (long)p));
    // S, D, obj = solve_s_graphical_group(Symmetric(sigma), groups, group_reps, method; m, kwargs...) (group.jl:192-193)
    std::vector<double> S11 = gather_sigma11(Sigma, p, reps);          // cond_indep_corr keeps Σ[reps, reps]
    double obj = 0.0;
    KB_TRY(solve_s_graphical_group_dev(ctx, dSig, (int)p, S11.data(), groups, reps, method, (double)m, opts, V, (int)nv, S11_out,
                                       dD, &obj, info));
    if (obj_out) *obj_out = obj;
    if (D_out) KB_CUDA(ctx, cudaMemcpyAsync(D_out, dD, sizeof(double) * pp, cudaMemcpyDeviceToHost, ctx->stream));
    // X̃ = condition(X, μ, Σ, Symmetric(D); m)                                                      (group.jl:196)
    KB_TRY(sc.alloc(&dSiS, pp));
    KB_TRY(sc.alloc(&dLb, pp * nblk));
    KB_TRY(cond_factor_dev(ctx, dRaw, (long)p, (int)p, dD, (long)p, 0, m, dSiS, dLb, (long)p));
    return sample_host_pipeline(ctx, X, (long)n, (int)p, dmu, dSiS, dLb, (long)p, m, Z, seed, 0, Xko_out);
}

// ---- test hooks -------------------------------------------------------------------------------------
int kb_test_dgemm(kb_ctx* ctx, int32_t transa, int32_t transb, int64_t M, int64_t N, int64_t K, double alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const long ar = transa ? K : M, ac = transa ? M : K, br = transb ? N : K, bc = transb ? K : N;
    double *dA, *dB, *dC;
    const long dlda = (ar + 1) & ~1L, dldb = (br + 1) & ~1L, dldc = (M + 1) & ~1L;
    KB_TRY(sc.alloc(&dA, (size_t)dlda * ac));
    KB_TRY(sc.alloc(&dB, (size_t)dldb * bc));
    KB_TRY(sc.alloc(&dC, (size_t)dldc * N));
    KB_TRY(h2d_matrix(ctx, A, lda, dA, dlda, ar, ac, ctx->stream));
    KB_TRY(h2d_matrix(ctx, B, ldb, dB, dldb, br, bc, ctx->stream));
    KB_TRY(h2d_matrix(ctx, C, ldc, dC, dldc, M, N, ctx->stream));
    KB_TRY(gemm1(ctx, transa ? A_KCONTIG : A_MCONTIG, transb ? B_NCONTIG : B_KCONTIG, (int)M, (int)N, (int)K, alpha, dA,
                 dlda, dB, dldb, beta, dC, dldc));
    KB_TRY(d2h_matrix(ctx, dC, dldc, C, ldc, M, N, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_test_dgemm_dev(kb_ctx* ctx, int32_t transa, int32_t transb, int64_t M, int64_t N, int64_t K, double alpha,
                      const double* dA, int64_t lda, const double* dB, int64_t ldb, double beta, double* dC, int64_t ldc) {
    if (!ctx) return bad_ctx();
    if (!dA || !dB || !dC || M < 1 || N < 1 || K < 1) return set_error(ctx, KB_ERR_INVALID, "test_dgemm_dev: bad arguments");
    DeviceGuard g(ctx->device);
    return gemm1(ctx, transa ? A_KCONTIG : A_MCONTIG, transb ? B_NCONTIG : B_KCONTIG, (int)M, (int)N, (int)K, alpha, dA, (long)lda,
                 dB, (long)ldb, beta, dC, (long)ldc);
}

int kb_test_factor_band(kb_ctx* ctx, const double* Lblocks, int64_t p, int32_t m, int32_t* band_out) {
    if (!ctx) return bad_ctx();
    if (!Lblocks || !band_out || p < 1 || m < 1) return set_error(ctx, KB_ERR_INVALID, "test_factor_band: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const size_t cnt = (size_t)p * p * (2 * m - 1);
    double* dLb;
    KB_TRY(sc.alloc(&dLb, cnt));
    KB_CUDA(ctx, cudaMemcpyAsync(dLb, Lblocks, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->stream));
    int band = 0;
    KB_TRY(factor_band_dev(ctx, dLb, (long)p, (int)p, m, &band));
    *band_out = band;
    return KB_OK;
}

int kb_test_potrf_positive(kb_ctx* ctx, const double* A, int64_t p, double* L_out, double* d_out, int32_t* modified_out) {
    if (!ctx) return bad_ctx();
    if (!A || !L_out || !d_out || p < 1) return set_error(ctx, KB_ERR_INVALID, "test_potrf_positive: bad arguments");
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    const long ld = (p + 1) & ~1L;
    double *dA, *dL, *dd;
    KB_TRY(sc.alloc(&dA, (size_t)p * p));
    KB_TRY(sc.alloc(&dL, (size_t)ld * p));
    KB_TRY(sc.alloc(&dd, (size_t)p));
    KB_CUDA(ctx, cudaMemcpyAsync(dA, A, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(sym_upper_to_full_dev(ctx, dA, (long)p, (int)p, dL, ld));      // Symmetric(A), uplo = :U
    int modified = 0;
    KB_TRY(potrf_positive_lower(ctx, dL, ld, (int)p, dd, &modified));
    KB_TRY(d2h_matrix(ctx, dL, ld, L_out, (long)p, (long)p, (long)p, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(d_out, dd, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (modified_out) *modified_out = modified;
    return KB_OK;
}

int kb_test_potrf(kb_ctx* ctx, const double* A, int64_t p, double* L_out) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dA, *dL;
    KB_TRY(sc.alloc(&dA, (size_t)p * p));
    KB_TRY(sc.alloc(&dL, (size_t)p * p));
    KB_CUDA(ctx, cudaMemcpyAsync(dA, A, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(fill_zero(ctx, dL, (size_t)p * p));
    KB_TRY(potrf_dev(ctx, dA, (long)p, (int)p, dL, (long)p));
    KB_CUDA(ctx, cudaMemcpyAsync(L_out, dL, sizeof(double) * p * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_test_spd_inverse(kb_ctx* ctx, const double* A, int64_t p, double* Ainv_out) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double *dA, *dO;
    KB_TRY(sc.alloc(&dA, (size_t)p * p));
    KB_TRY(sc.alloc(&dO, (size_t)p * p));
    KB_CUDA(ctx, cudaMemcpyAsync(dA, A, sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_TRY(fill_zero(ctx, dO, (size_t)p * p));
    KB_TRY(spd_inverse_dev(ctx, dA, (long)p, (int)p, dO, (long)p));
    KB_CUDA(ctx, cudaMemcpyAsync(Ainv_out, dO, sizeof(double) * p * p, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

int kb_test_randn(kb_ctx* ctx, uint64_t seed, int64_t row_offset, int64_t n, int64_t ncols, double* Z_out) {
    if (!ctx) return bad_ctx();
    DeviceGuard g(ctx->device);
    Scope sc(ctx);
    double* dZ;
    KB_TRY(sc.alloc(&dZ, (size_t)n * ncols));
    KB_TRY(randn_dev(ctx, seed, row_offset, (long)n, (long)ncols, 0, dZ, (long)n));
    KB_CUDA(ctx, cudaMemcpyAsync(Z_out, dZ, sizeof(double) * n * ncols, cudaMemcpyDeviceToHost, ctx->stream));
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

}  // extern "C"
