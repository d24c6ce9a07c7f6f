// Multi-GPU behind the C ABI (SURVEY 8b `kb_create(ctx**, const int* devs, int ndev)`, 8e): ONE host process, one
// kb_ctx + one host thread per device.  Only the naturally data-parallel pieces are partitioned:
//   * Gram: rows of X sharded, per-device DMMA SYRK, then ONE all-reduce of the p x p sums written here as a single
//     peer-to-peer kernel per device (device d sums slice d of every peer's buffer over NVLink in a fixed order and
//     stores the result back into every peer: bit-identical on all devices, no NCCL, no host staging);
//   * condition / sampling: rows sharded, no exchange (the RNG is keyed by the global row);
//   * independent solves ("replicas only": a single CCD solve is sequential) dealt round-robin to the devices;
//   * group solve of a block-diagonal Σ: LD blocks on different devices sweep in lockstep, the reference's global
//     decisions (shared equi γ, convergence on the summed objective / max δ) go through an in-process reduce.
#include "kb_common.cuh"
#include "linalg.cuh"
#include "pipeline.cuh"
#include "group.cuh"
#include "hostpipe.cuh"
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>

using namespace kb;

constexpr int KB_MAX_DEVICES = 16;

struct kb_multi {
    std::vector<kb_ctx*> ctx;
    std::string last_error;
    bool peer_ok = true;
};

namespace {

struct PeerPtrs {
    double* p[KB_MAX_DEVICES];
    int n;
};

// all-reduce(sum) of `count` doubles held in n peer buffers, slice [e0, e1) handled by this launch: peer loads in a
// FIXED device order (every device ends with the same bits), peer stores of the result into every buffer.
__global__ void p2p_allreduce_slice_kernel(PeerPtrs bufs, long e0, long e1) {
    const long stride = (long)gridDim.x * blockDim.x * 2;
    for (long e = e0 + 2 * ((long)blockIdx.x * blockDim.x + threadIdx.x); e < e1; e += stride) {
        if (e + 1 < e1) {
            double2 v = make_double2(0.0, 0.0);
            for (int g = 0; g < bufs.n; ++g) {
                const double2 x = *reinterpret_cast<const double2*>(bufs.p[g] + e);
                v.x += x.x; v.y += x.y;
            }
            for (int g = 0; g < bufs.n; ++g) *reinterpret_cast<double2*>(bufs.p[g] + e) = v;
        } else {
            double v = 0.0;
            for (int g = 0; g < bufs.n; ++g) v += bufs.p[g][e];
            for (int g = 0; g < bufs.n; ++g) bufs.p[g][e] = v;
        }
    }
}

struct ThreadBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, count = 0, gen = 0;
    bool aborted = false;
    explicit ThreadBarrier(int n_) : n(n_) {}
    // returns false if some thread aborted
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return false;
        const int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g || aborted; });
        return !aborted;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};

// Runs fn(i) on one host thread per context; returns the first non-OK status (its message goes to mc->last_error).
int run_on_all(kb_multi* mc, int nthreads, const std::function<int(int)>& fn, const std::function<kb_ctx*(int)>& ctx_of) {
    std::vector<int> rc((size_t)nthreads, KB_OK);
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i)
        th.emplace_back([&, i] {
            kb_ctx* c = ctx_of(i);
            cudaSetDevice(c->device);
            rc[(size_t)i] = fn(i);
        });
    for (auto& t : th) t.join();
    // report the root cause, not the "another device failed" echo of the threads that were waiting for it
    int first = -1;
    for (int i = 0; i < nthreads; ++i)
        if (rc[(size_t)i] != KB_OK) {
            if (first < 0) first = i;
            if (ctx_of(i)->last_error.find("another device failed") == std::string::npos &&
                ctx_of(i)->last_error.find("reduce hook failed") == std::string::npos) { first = i; break; }
        }
    if (first < 0) return KB_OK;
    mc->last_error = "device thread " + std::to_string(first) + ": " + ctx_of(first)->last_error;
    return rc[(size_t)first];
}

int merr(kb_multi* mc, int code, const char* msg) {
    if (mc) mc->last_error = msg;
    return code;
}

void shard_rows(long n, int world, int rank, long* lo, long* hi) {
    const long base = n / world, rem = n % world;
    *lo = rank * base + std::min<long>(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// in-process reduce of a few doubles between the host threads of a lockstep group solve (kb_reduce_fn)
struct Rendezvous {
    std::mutex mu;
    std::condition_variable cv;
    int n, count = 0, gen = 0;
    bool aborted = false;
    std::vector<std::vector<double>> slot;
    std::vector<double> result;
    explicit Rendezvous(int n_) : n(n_), slot((size_t)n_) {}
};
struct RendezvousUser {
    Rendezvous* r;
    int id;
};
int rendezvous_reduce(void* user, int32_t op, double* vals, int32_t nv) {
    auto* u = static_cast<RendezvousUser*>(user);
    Rendezvous& r = *u->r;
    std::unique_lock<std::mutex> lk(r.mu);
    if (r.aborted) return 1;
    const int g = r.gen;
    r.slot[(size_t)u->id].assign(vals, vals + nv);
    if (++r.count == r.n) {
        // combine in block order: the result does not depend on which thread arrived last
        r.result = r.slot[0];
        for (int b = 1; b < r.n; ++b)
            for (int k = 0; k < nv; ++k) {
                const double x = r.slot[(size_t)b][(size_t)k];
                double& y = r.result[(size_t)k];
                y = (op == KB_REDUCE_MIN) ? std::min(y, x) : (op == KB_REDUCE_MAX) ? std::max(y, x) : y + x;
            }
        r.count = 0;
        ++r.gen;
        r.cv.notify_all();
    } else {
        r.cv.wait(lk, [&] { return r.gen != g || r.aborted; });
        if (r.aborted) return 1;
    }
    std::copy(r.result.begin(), r.result.begin() + nv, vals);
    return 0;
}

}  // namespace

extern "C" {

int kb_create_multi(kb_multi** out, const int* devices, int32_t ndev) {
    if (!out) return KB_ERR_INVALID;
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > KB_MAX_DEVICES) return KB_ERR_INVALID;
    kb_multi* mc = new (std::nothrow) kb_multi();
    if (!mc) return KB_ERR_CUDA;
    for (int i = 0; i < ndev; ++i) {
        kb_ctx* c = nullptr;
        const int rc = kb_create(&c, devices[i]);
        if (rc != KB_OK) {
            for (auto* x : mc->ctx) kb_destroy(x);
            delete mc;
            return rc;
        }
        mc->ctx.push_back(c);
    }
    // peer access between every pair of distinct devices (NVLink / NVSwitch on the HGX board)
    int prev = 0;
    cudaGetDevice(&prev);
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < ndev; ++j) {
            const int a = devices[i], b = devices[j];
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (!can) { mc->peer_ok = false; continue; }
            cudaSetDevice(a);
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mc->peer_ok = false;
            cudaGetLastError();
        }
    cudaSetDevice(prev);
    *out = mc;
    return KB_OK;
}

void kb_destroy_multi(kb_multi* mc) {
    if (!mc) return;
    for (auto* c : mc->ctx) kb_destroy(c);
    delete mc;
}

int32_t kb_multi_size(const kb_multi* mc) { return mc ? (int32_t)mc->ctx.size() : 0; }
kb_ctx* kb_multi_ctx(kb_multi* mc, int32_t i) { return (mc && i >= 0 && i < (int32_t)mc->ctx.size()) ? mc->ctx[(size_t)i] : nullptr; }
const char* kb_multi_last_error(const kb_multi* mc) { return mc ? mc->last_error.c_str() : "invalid multi-context"; }

// Σ̂ = (Xc'Xc)/denom and the column means, rows of X sharded over the devices, one peer-to-peer all-reduce.
int kb_multi_gram(kb_multi* mc, const double* X, int64_t n, int64_t p, int32_t center, double denom, double* G_out,
                  double* colmean_out) {
    if (!mc) return KB_ERR_INVALID;
    if (!X || !G_out || p < 1 || n < 1) return merr(mc, KB_ERR_INVALID, "multi_gram: bad arguments");
    const int nd = (int)mc->ctx.size();
    if (nd > 1 && !mc->peer_ok) return merr(mc, KB_ERR_CUDA, "multi_gram: peer access between the devices is not available");
    const long ld = (p + 1) & ~1L;
    const long count = ld * p;
    PeerPtrs G{}, CS{};
    G.n = CS.n = nd;
    ThreadBarrier bar(nd);
    auto fn = [&](int i) -> int {
        kb_ctx* ctx = mc->ctx[(size_t)i];
        Scope sc(ctx);
        double *dG = nullptr, *dcs = nullptr;
        int rc = sc.alloc(&dG, (size_t)count);
        if (rc == KB_OK) rc = sc.alloc(&dcs, (size_t)((p + 1) & ~1L));
        if (rc == KB_OK) rc = fill_zero(ctx, dG, (size_t)count);
        if (rc == KB_OK) rc = fill_zero(ctx, dcs, (size_t)((p + 1) & ~1L));
        long lo = 0, hi = 0;
        shard_rows((long)n, nd, i, &lo, &hi);
        if (rc == KB_OK) rc = gram_accumulate_host_rows(ctx, X + lo, hi - lo, (long)n, (int)p, dG, ld, dcs);
        if (rc == KB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, KB_ERR_CUDA, "multi_gram: sync failed");
        G.p[i] = dG; CS.p[i] = dcs;
        if (rc != KB_OK) { bar.abort(); return rc; }
        if (!bar.wait()) return set_error(ctx, KB_ERR_CUDA, "multi_gram: another device failed");
        if (nd > 1) {
            // slice i of the p x p sums (and of the column sums): peer loads + peer stores over NVLink
            const long per = ((count / nd) + 1) & ~1L;
            const long e0 = std::min(count, per * i), e1 = (i == nd - 1) ? count : std::min(count, per * (i + 1));
            static const bool trace = getenv("KB_MULTI_TRACE") != nullptr;
            if (trace) cudaEventRecord(ctx->ev0, ctx->stream);
            if (e1 > e0) {
                p2p_allreduce_slice_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(G, e0, e1);
                ctx->launches++;
            }
            if (trace) {
                cudaEventRecord(ctx->ev1, ctx->stream);
                cudaEventSynchronize(ctx->ev1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
                // each device reads its slice from nd buffers and writes it to nd buffers
                fprintf(stderr, "[multi_gram dev %d] p2p all-reduce slice: %ld doubles x %d peers, %.3f ms, %.1f GB/s read + %.1f GB/s written\n",
                        ctx->device, e1 - e0, nd, ms, (e1 - e0) * 8.0 * nd / ms / 1e6, (e1 - e0) * 8.0 * nd / ms / 1e6);
            }
            if (i == 0) {
                p2p_allreduce_slice_kernel<<<8, 256, 0, ctx->stream>>>(CS, 0, (long)p);
                ctx->launches++;
            }
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
                bar.abort();
                return set_error(ctx, KB_ERR_CUDA, "multi_gram: peer-to-peer all-reduce failed");
            }
            if (!bar.wait()) return set_error(ctx, KB_ERR_CUDA, "multi_gram: another device failed");
        }
        // every device now holds the full sums; device 0 finishes and returns them
        if (i == 0) {
            rc = gram_finish_dev(ctx, dG, ld, dcs, (long)n, (int)p, center, denom);
            if (rc == KB_OK) rc = d2h_matrix(ctx, dG, ld, G_out, (long)p, (long)p, (long)p, ctx->stream);
            if (rc == KB_OK && colmean_out &&
                cudaMemcpyAsync(colmean_out, dcs, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                rc = set_error(ctx, KB_ERR_CUDA, "multi_gram: copy-out failed");
            if (rc == KB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_error(ctx, KB_ERR_CUDA, "multi_gram: sync failed");
        }
        // peers may still be reading this device's buffers until everyone has passed the second barrier (done above)
        return rc;
    };
    return run_on_all(mc, nd, fn, [&](int i) { return mc->ctx[(size_t)i]; });
}

// condition(X, μ, Σ, S; m) with the rows of X (and of Z / X̃) sharded over the devices; every device builds the factor.
int kb_multi_condition(kb_multi* mc, const double* X, int64_t n, int64_t p, const double* mu, const double* Sigma,
                       const double* S, int32_t s_is_diag, int32_t m, const double* Z, uint64_t seed, double* Xko_out) {
    if (!mc) return KB_ERR_INVALID;
    if (!X || !mu || !Sigma || !S || !Xko_out || p < 1 || n < 0) return merr(mc, KB_ERR_INVALID, "multi_condition: bad arguments");
    if (m < 1) return merr(mc, KB_ERR_INVALID, "m should be 1 or larger.");
    const int nd = (int)mc->ctx.size();
    auto fn = [&](int i) -> int {
        long lo = 0, hi = 0;
        shard_rows((long)n, nd, i, &lo, &hi);
        if (hi <= lo) return KB_OK;
        return condition_host(mc->ctx[(size_t)i], X + lo, hi - lo, (long)n, (int)p, mu, Sigma, S, s_is_diag, m, Z ? Z + lo : nullptr, seed,
                              lo, Xko_out + lo);
    };
    return run_on_all(mc, nd, fn, [&](int i) { return mc->ctx[(size_t)i]; });
}

// modelX_gaussian_knockoffs(X, method, μ, Σ; m) (modelX.jl:53-66): the solve on device 0 (sequential sweep), then the
// rows of X sharded over all devices for `condition`.
int kb_multi_modelx_gaussian_knockoffs(kb_multi* mc, const double* X, int64_t n, int64_t p, const double* mu, const double* Sigma,
                                       int32_t method, int32_t m, const kb_solve_opts* opts, const double* s_init, const double* Z,
                                       uint64_t seed, double* Xko_out, double* s_out, kb_solve_info* info) {
    if (!mc) return KB_ERR_INVALID;
    if (!X || !mu || !Sigma || !Xko_out || !s_out || p < 1 || n < 0) return merr(mc, KB_ERR_INVALID, "multi modelX: bad arguments");
    if (m < 1) return merr(mc, KB_ERR_INVALID, "m should be 1 or larger.");
    kb_ctx* c0 = mc->ctx[0];
    const int rc = kb_solve_s(c0, Sigma, p, method, (double)m, opts, s_init, s_out, info);
    if (rc != KB_OK) { mc->last_error = c0->last_error; return rc; }
    return kb_multi_condition(mc, X, n, p, mu, Sigma, s_out, 1, m, Z, seed, Xko_out);
}

// "replicas only": nprob independent solve_s problems on the SAME Σ (different methods / m) dealt round-robin to the devices
// (BASELINE configs[2]: :sdp_ccd and :mvr side by side).  s_out: p x nprob column-major; infos: nprob entries (may be NULL).
int kb_multi_solve_s_many(kb_multi* mc, const double* Sigma, int64_t p, int32_t nprob, const int32_t* methods, const double* ms,
                          const kb_solve_opts* opts, double* s_out, kb_solve_info* infos) {
    if (!mc) return KB_ERR_INVALID;
    if (!Sigma || !methods || !ms || !s_out || p < 1 || nprob < 1) return merr(mc, KB_ERR_INVALID, "multi solve_s_many: bad arguments");
    const int nd = std::min((int)mc->ctx.size(), (int)nprob);
    auto fn = [&](int i) -> int {
        for (int q = i; q < nprob; q += nd) {
            const int rc = kb_solve_s(mc->ctx[(size_t)i], Sigma, p, methods[q], ms[q], opts, nullptr, s_out + (size_t)q * p,
                                      infos ? infos + q : nullptr);
            if (rc != KB_OK) return rc;
        }
        return KB_OK;
    };
    return run_on_all(mc, nd, fn, [&](int i) { return mc->ctx[(size_t)i]; });
}

// solve_s_group on Σ = blockdiag(Σ_0 .. Σ_{B-1}) (no group straddles blocks; SURVEY 8e, BASELINE configs[3]): block b on device
// b mod ndev, all blocks sweep in lockstep, global decisions reduced in-process.  Sigma_blocks[b]: p_b x p_b; groups[b]: p_b ids;
// S_out[b]: p_b x p_b.  obj_out = the summed objective; infos: B entries (may be NULL).
int kb_multi_solve_s_group_blockdiag(kb_multi* mc, int32_t nblocks, const double* const* Sigma_blocks, const int64_t* p_blocks,
                                     const int64_t* const* groups, int32_t method, double m, const kb_group_opts* opts,
                                     double* const* S_out, double* obj_out, kb_group_info* infos) {
    if (!mc) return KB_ERR_INVALID;
    if (nblocks < 1 || nblocks > 1024 || !Sigma_blocks || !p_blocks || !groups || !S_out)
        return merr(mc, KB_ERR_INVALID, "multi group solve: bad arguments");
    const int nd = (int)mc->ctx.size();
    // one context per block: the first ndev blocks use the multi-context's own, the others get temporary ones
    std::vector<kb_ctx*> bctx((size_t)nblocks, nullptr);
    std::vector<kb_ctx*> temp;
    for (int b = 0; b < nblocks; ++b) {
        if (b < nd) { bctx[(size_t)b] = mc->ctx[(size_t)b]; continue; }
        kb_ctx* c = nullptr;
        const int rc = kb_create(&c, mc->ctx[(size_t)(b % nd)]->device);
        if (rc != KB_OK) {
            for (auto* t : temp) kb_destroy(t);
            return merr(mc, rc, "multi group solve: could not create a per-block context");
        }
        temp.push_back(c);
        bctx[(size_t)b] = c;
    }
    Rendezvous rv(nblocks);
    std::vector<RendezvousUser> users((size_t)nblocks);
    std::vector<double> objs((size_t)nblocks, 0.0);
    auto fn = [&](int b) -> int {
        users[(size_t)b] = RendezvousUser{&rv, b};
        const int rc = solve_s_group_host(bctx[(size_t)b], Sigma_blocks[b], (int)p_blocks[b], groups[b], method, m, opts, nullptr, 0,
                                          nblocks > 1 ? rendezvous_reduce : nullptr, &users[(size_t)b], S_out[b], &objs[(size_t)b],
                                          infos ? infos + b : nullptr);
        if (rc != KB_OK) {
            std::lock_guard<std::mutex> lk(rv.mu);
            rv.aborted = true;
            rv.cv.notify_all();
        }
        return rc;
    };
    const int rc = run_on_all(mc, nblocks, fn, [&](int b) { return bctx[(size_t)b]; });
    for (auto* t : temp) kb_destroy(t);
    if (rc != KB_OK) return rc;
    if (obj_out) {
        double t = 0.0;
        for (double o : objs) t += o;
        *obj_out = t;
    }
    return KB_OK;
}

}  // extern "C"
