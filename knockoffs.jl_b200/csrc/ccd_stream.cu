// Coordinate-descent s-solvers (MVR / maximum entropy / SDP-ccd) as ONE persistent kernel per launch.
//
// Reference (paths relative to biona001/Knockoffs.jl v2.0.2): solve_MVR src/utilities.jl:124-175,
// solve_max_entropy :221-280, solve_sdp_ccd :291-343, rank-1 factor updates :516-625.
//
// Device formulation.  The reference keeps the Cholesky factor U of A = κΣ − S (+ λmin I) and per
// coordinate j does 2-3 triangular solves with e_j / κΣ[:,j] and one Givens rank-1 up/downdate --
// every one of them an O(p)-deep dependency chain.  All the scalars those solves produce are
// entries of A⁻¹:   ‖L⁻¹e_j‖² = (A⁻¹)_jj,   ‖A⁻¹e_j‖² = ‖(A⁻¹)[:,j]‖²,
// ‖L⁻¹ỹ‖² = A_jj (A_jj (A⁻¹)_jj − 1)  for ỹ = κΣ[:,j] with ỹ_j = 0 (because ỹ = A e_j − A_jj e_j).
// So the kernel keeps the symmetric inverse M = A⁻¹ (lower triangle, column-major) resident in HBM
// and the factor modification A ← A − δ e_j e_jᵀ becomes the rank-1 Sherman–Morrison update
//      M ← M + c · m_j m_jᵀ,   c = δ / (1 − δ M_jj),   m_j = M[:, j]
// -- a perfectly parallel, fully coalesced streaming pass over the p(p+1)/2 stored entries
// (8 B read + 8 B written per entry = 8p² B per coordinate, the same count as the reference's
// solves + update), with no dependency chain at all.  The only cross-CTA communication
// per coordinate is the next column m_{j+1}, which its owners publish (already updated) BEFORE
// they stream the rest of their slots, so the flag wait at the next coordinate is normally
// already satisfied.
//
// Work decomposition: the lower triangle is cut into "slots" = (column k, 64-row block rb >= k/64);
// slots are enumerated column-major and each CTA owns a contiguous, equal-sized slot range.  The
// buffer's leading dimension is a multiple of 64 with zero padding rows, and the 64 x 64 diagonal
// blocks are kept fully symmetric, so every slot is one unmasked 512-byte warp access
// (32 lanes x double2) -- no per-element predicates in the streaming loop.
// Line-search scalars (‖m_j‖², max|m| per 64-row block) use warp-shuffle reductions; every CTA
// derives bit-identical scalars from the same published column, so all CTAs take identical
// skip / clip / convergence decisions without any further exchange.
#include "kb_common.cuh"
#include "ccd.cuh"
#include "ccd_device.cuh"
#include <cstdlib>

namespace kb {

template <int CCD_THREADS, int MIN_CTAS, int CCD_ILP, bool DOUBLE_BUF>
__global__ void __launch_bounds__(CCD_THREADS, MIN_CTAS) ccd_stream_kernel(const CcdParams P) {
    constexpr int CCD_WARPS = CCD_THREADS / 32;
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_abort;
    const int p = P.p;
    const long ld = P.ld;                  // multiple of 64
    const int nrb = (int)(ld / SLOT_ROWS);
    double* col = smem;                    // current column m_j, ld doubles (rows >= p are zero)
    double* rbmax = smem + ld;             // max|m| per 64-row block
    double* red = rbmax + ((nrb + 3) & ~3);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* __restrict__ M = P.M;
    const long colstride = ld + 8;         // colbuf entry: ld values (rows >= p zero), then s_j, adiag_j

    const long total_slots = slot_index(p - 1, nrb - 1, nrb) + 1;
    const long s_begin = total_slots * blockIdx.x / gridDim.x;
    const long s_end = total_slots * (blockIdx.x + 1) / gridDim.x;
    const bool have_slots = s_end > s_begin;
    const SlotPos first = slot_from_index(have_slots ? s_begin : 0, nrb);
    const SlotPos last = slot_from_index(have_slots ? s_end - 1 : 0, nrb);

    // slot table: (element offset of the slot's first row in column k, (k << 16) | rb)
    int2* slots = reinterpret_cast<int2*>(red + CCD_WARPS + 4);
    const int nslots = (int)(s_end - s_begin);
    {
        int k = first.k, rb = first.rb;
        // every thread walks to its own entries (one linear walk of the CTA's range, done once per launch)
        for (int i = 0; i < nslots; ++i) {
            if ((i % CCD_THREADS) == tid) slots[i] = make_int2((int)((long)rb * SLOT_ROWS + (long)k * ld), (k << 16) | rb);
            if (++rb >= nrb) { ++k; rb = k / SLOT_ROWS; }
        }
    }
    double* const Mlane = M + 2 * lane;
    __syncthreads();

    const double m = P.m;
    const int method = P.method;
    double lam = P.sdp_lambda;
    unsigned long long target = gridDim.x;
    int buf = 0;
    int status = 0;
    int sweeps = 0;
    double touched = 0.0, nupd = 0.0;
    long long t_wait = 0, t_stage = 0, t_pub = 0, t_stream = 0, t_mark = clock64();   // phase cycle counters (thread 0)
#define KB_PHASE(acc)                                   \
    do {                                                \
        const long long _t = clock64();                 \
        acc += _t - t_mark;                             \
        t_mark = _t;                                    \
    } while (0)

    for (int l = 0; l < P.nsweeps && status == 0; ++l) {
        ++sweeps;
        double max_delta = 0.0;
        for (int j = 0; j < p; ++j) {
            // ---- 1. wait for column j (published during the previous coordinate) ----
            if (tid == 0) {
                int ab = 0;
                unsigned spins = 0;
                const long long t0 = clock64();
                while (ld_acquire_u64(P.ready) < target) {
                    if ((++spins & 255u) == 0u) {
                        if (ld_acquire_u64(P.ready + 1) != 0ULL) { ab = 1; break; }
                        if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
                            atomicExch(P.ready + 1, 1ULL);
                            ab = 1;
                            break;
                        }
                    }
                }
                s_abort = ab;
            }
            __syncthreads();
            if (s_abort) { status = KB_ERR_CUDA; break; }
            KB_PHASE(t_wait);
            target += gridDim.x;
            const double* cb = P.colbuf + (long)buf * colstride;
            // ---- 2. stage the column, sum of squares, per-row-block max ----
            double ss = 0.0;
            for (int rb0 = warp; rb0 < nrb; rb0 += 4 * CCD_WARPS) {
                // 4 independent L2 loads in flight per lane before any of them is consumed
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rb = rb0 + u * CCD_WARPS;
                    v[u] = make_double2(0.0, 0.0);
                    if (rb < nrb) v[u] = __ldcg(reinterpret_cast<const double2*>(cb + rb * SLOT_ROWS + 2 * lane));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rb = rb0 + u * CCD_WARPS;
                    if (rb < nrb) {
                        *reinterpret_cast<double2*>(col + rb * SLOT_ROWS + 2 * lane) = v[u];
                        ss = fma(v[u].x, v[u].x, ss);
                        ss = fma(v[u].y, v[u].y, ss);
                        double mx = fmax(fabs(v[u].x), fabs(v[u].y));
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        if (lane == 0) rbmax[rb] = mx;
                    }
                }
            }
            const double sj = __ldcg(cb + ld);
            const double ajj = __ldcg(cb + ld + 1);
            ss = block_sum<CCD_WARPS>(ss, red, tid);          // includes the barrier that publishes col[] / rbmax[]
            const double Mjj = col[j];

            // ---- 3. line search (every thread, identical arithmetic) ----
            double delta, s_new, cfac;
            bool skip;
            ccd_line_search(method, m, lam, P.lambda_min, sj, ajj, (method == KB_METHOD_MVR) ? 0.0 : P.ksig_diag[j], ss, Mjj,
                            delta, s_new, skip, cfac, status);
            if (!skip && status == 0) {
                if (fabs(delta) > max_delta) max_delta = fabs(delta);
                nupd += 1.0;
            }

            KB_PHASE(t_stage);
            // ---- 4. publish column jn = j+1 (already updated) ----
            const int jn = (j + 1 == p) ? 0 : j + 1;
            const int nbuf = (buf + 1 == 3) ? 0 : buf + 1;
            double* ob = P.colbuf + (long)nbuf * colstride;
            const int rbn = jn / SLOT_ROWS;
            if (have_slots) {
                const double mjn = col[jn];
                // (a) row jn of my columns k < jn:   entry (jn, k) lives in slot (k, rbn)
                const int kend = min(last.k, jn - 1);
                for (int k = first.k + tid; k <= kend; k += CCD_THREADS) {
                    const long si = slot_index(k, rbn, nrb);
                    if (si >= s_begin && si < s_end) {
                        const double old = M[(long)jn + (long)k * ld];
                        ob[k] = fma(cfac * col[k], mjn, old);
                    }
                }
                // (b) column jn itself, rows i >= jn, for the slots (jn, rb) that are mine
                if (jn >= first.k && jn <= last.k) {
                    const double cjn = cfac * mjn;
                    const int rb_lo = (jn == first.k) ? max(first.rb, rbn) : rbn;
                    const int rb_hi = (jn == last.k) ? last.rb : nrb - 1;
                    const int i_lo = max(jn, rb_lo * SLOT_ROWS);
                    const int i_hi = min(p, (rb_hi + 1) * SLOT_ROWS);
                    for (int i = i_lo + tid; i < i_hi; i += CCD_THREADS) {
                        const double old = M[(long)i + (long)jn * ld];
                        ob[i] = fma(cjn, col[i], old);
                    }
                }
            }
            // s_j / A_jj bookkeeping: the owner of the diagonal slot of column j records them; the
            // owner of column jn's diagonal slot forwards s_jn, A_jn,jn with the column.
            {
                const long sdiag_j = slot_index(j, j / SLOT_ROWS, nrb);
                if (tid == 0 && sdiag_j >= s_begin && sdiag_j < s_end && !skip && status == 0) {
                    P.s[j] = s_new;
                    P.adiag[j] = ajj - delta;
                }
                __syncthreads();   // s[j] visible to this CTA's thread 0 below if jn == j (p == 1)
                const long sdiag_n = slot_index(jn, rbn, nrb);
                if (tid == 0 && sdiag_n >= s_begin && sdiag_n < s_end) {
                    ob[ld] = P.s[jn];
                    ob[ld + 1] = P.adiag[jn];
                }
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(P.ready, 1ULL);
            }
            buf = nbuf;
            KB_PHASE(t_pub);

            // ---- 5. stream my slots:  M(i,k) += (c m_k) m_i ----
            // Slots come from the CTA's shared-memory table (element offset, column, row block); each warp
            // takes slots warp, warp+W, ... in batches of CCD_ILP.  The loads of batch b+1 are issued before
            // the FMAs / stores of batch b (slots never alias), so 2 x CCD_ILP 512-byte requests are in
            // flight per warp.
            if (cfac != 0.0 && nslots > 0) {
                double2 va[CCD_ILP], vb[CCD_ILP];
                int2 ea[CCD_ILP], eb[CCD_ILP];
                double cka[CCD_ILP], ckb[CCD_ILP];
                int nact = 0;
                auto load_batch = [&](int i0, double2 (&v)[CCD_ILP], int2 (&e)[CCD_ILP], double (&ck)[CCD_ILP]) {
#pragma unroll
                    for (int u = 0; u < CCD_ILP; ++u) {
                        const int i = i0 + u * CCD_WARPS;
                        ck[u] = 0.0;
                        if (i < nslots) {
                            e[u] = slots[i];
                            const double c = cfac * col[e[u].y >> 16];
                            if (fabs(c) * rbmax[e[u].y & 0xffff] >= SKIP_ABS) {
                                ck[u] = c;
                                v[u] = *reinterpret_cast<const double2*>(Mlane + e[u].x);
                            }
                        }
                    }
                };
                auto store_batch = [&](double2 (&v)[CCD_ILP], int2 (&e)[CCD_ILP], double (&ck)[CCD_ILP]) {
#pragma unroll
                    for (int u = 0; u < CCD_ILP; ++u) {
                        if (ck[u] != 0.0) {
                            const double2 mv = *reinterpret_cast<const double2*>(col + (e[u].y & 0xffff) * SLOT_ROWS + 2 * lane);
                            v[u].x = fma(ck[u], mv.x, v[u].x);
                            v[u].y = fma(ck[u], mv.y, v[u].y);
                            *reinterpret_cast<double2*>(Mlane + e[u].x) = v[u];
                            ++nact;
                        }
                    }
                };
                int i0 = warp;
                if (DOUBLE_BUF) {
                    load_batch(i0, va, ea, cka);
                    while (true) {
                        const int i1 = i0 + CCD_ILP * CCD_WARPS;
                        if (i1 >= nslots) { store_batch(va, ea, cka); break; }
                        load_batch(i1, vb, eb, ckb);
                        store_batch(va, ea, cka);
                        i0 = i1 + CCD_ILP * CCD_WARPS;
                        if (i0 >= nslots) { store_batch(vb, eb, ckb); break; }
                        load_batch(i0, va, ea, cka);
                        store_batch(vb, eb, ckb);
                    }
                } else {
                    for (; i0 < nslots; i0 += CCD_ILP * CCD_WARPS) {
                        load_batch(i0, va, ea, cka);
                        store_batch(va, ea, cka);
                    }
                }
                if (lane == 0) touched += (double)nact;
            }
            __syncwarp();
            if (tid == 0) KB_PHASE(t_stream);
            if (status != 0) break;
        }
        if (blockIdx.x == 0 && tid == 0 && status == 0) {
            if (P.sweep0 + l < KB_MAX_TRACE) P.trace[P.sweep0 + l] = max_delta;
            P.counters[3] = max_delta;
        }
        if (status != 0) break;
        if (method == KB_METHOD_SDP_CCD) {
            lam *= P.sdp_mu;                       // utilities.jl:339-340
            if (lam < P.tol) break;
        } else if (max_delta < P.tol) {
            break;                                 // utilities.jl:172 / :277
        }
    }
    // per-CTA counters
    const double tsum = block_sum<CCD_WARPS>((lane == 0) ? touched : 0.0, red, tid);
    if (tid == 0) {
        atomicAdd(P.counters + 4, (double)t_wait);
        atomicAdd(P.counters + 5, (double)t_stage);
        atomicAdd(P.counters + 6, (double)t_pub);
        atomicAdd(P.counters + 7, (double)t_stream);
        atomicAdd(P.counters + 0, tsum);
        if (blockIdx.x == 0) {
            P.counters[1] += nupd;
            P.counters[2] = lam;
            *P.sweeps_out = sweeps;
            *P.status = status;
        }
    }
}

// column 0 of the current M (+ s_0, A_00) into colbuf[0]; ready = number of CTAs, abort flag cleared
__global__ void ccd_stream_init_kernel(const CcdParams P, unsigned long long nctas) {
    const int p = P.p;
    const long ld = P.ld;
    const long colstride = ld + 8;
    for (long i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * colstride; i += (long)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (i < p) v = P.M[i];              // column 0, rows 0..p-1 (all in the lower triangle)
        else if (i == ld) v = P.s[0];
        else if (i == ld + 1) v = P.adiag[0];
        P.colbuf[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.ready[0] = nctas;
        P.ready[1] = 0ULL;
        *P.status = 0;
        *P.sweeps_out = 0;
    }
}

size_t ccd_stream_colbuf_doubles(int p) { return 3 * (size_t)(ccd_stream_ld(p) + 8); }

template <int THREADS, int MIN_CTAS, int ILP, bool DB>
static int ccd_stream_launch(kb_ctx* ctx, const CcdParams& P, int* nctas_out) {
    constexpr int WARPS = THREADS / 32;
    auto kern = ccd_stream_kernel<THREADS, MIN_CTAS, ILP, DB>;
    const int p = P.p;
    const int nrb = (int)(P.ld / SLOT_ROWS);
    const long total_slots = slot_index(p - 1, nrb - 1, nrb) + 1;
    const size_t base_smem = (size_t)(P.ld + ((nrb + 3) & ~3) + WARPS + 4) * sizeof(double);
    // try MIN_CTAS CTAs per SM first, then fewer; the slot table (8 B per owned slot) must fit next to the column
    long nctas = 0;
    size_t smem = 0;
    for (int want = MIN_CTAS; want >= 1 && nctas == 0; --want) {
        long g = (long)ctx->num_sms * want;
        // small problems: keep a few slot passes per warp so CTAs are not empty
        const long min_slots_per_cta = 64;
        if (g > (total_slots + min_slots_per_cta - 1) / min_slots_per_cta)
            g = (total_slots + min_slots_per_cta - 1) / min_slots_per_cta;
        if (g < 1) g = 1;
        const long per_cta = (total_slots + g - 1) / g + 1;
        const size_t need = base_smem + (size_t)per_cta * sizeof(int2);
        if (need > 227 * 1024) continue;
        KB_CUDA(ctx, raise_max_dynamic_smem((const void*)kern, need));
        int per_sm = 0;
        KB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, need));
        if (per_sm >= want || g <= (long)ctx->num_sms * per_sm) { nctas = g; smem = need; }
    }
    if (nctas == 0) return set_error(ctx, KB_ERR_INVALID, "p = %d too large for the streaming CCD kernel", p);
    ccd_stream_init_kernel<<<8, 256, 0, ctx->stream>>>(P, (unsigned long long)nctas);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    // Optional: pin a fraction of the resident triangle in the 126 MB L2 (persisting access-policy window) so
    // the cyclic streaming pass does not evict what the next coordinate re-reads.  KB_CCD_L2_PERSIST = fraction
    // of the persisting carve-out to use (0 = off).
    static const double l2frac = [] {
        const char* e = getenv("KB_CCD_L2_PERSIST");
        return e ? atof(e) : 0.0;
    }();
    bool window = false;
    if (l2frac > 0.0 && ctx->l2_persist_max > 0 && ctx->l2_window_max > 0) {
        const size_t bytes = (size_t)P.ld * p * sizeof(double);
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = (void*)P.M;
        attr.accessPolicyWindow.num_bytes = bytes < ctx->l2_window_max ? bytes : ctx->l2_window_max;
        // roughly half of the window's address range is the unused upper triangle (never touched)
        double hr = l2frac * (double)ctx->l2_persist_max / (0.5 * (double)attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitRatio = (float)(hr > 1.0 ? 1.0 : hr);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        window = cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
    }
    void* args[] = {(void*)&P};
    KB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)nctas), dim3(THREADS), args, smem, ctx->stream));
    ctx->launches++;
    if (window) {
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    if (nctas_out) *nctas_out = (int)nctas;
    return KB_OK;
}

int ccd_stream_run(kb_ctx* ctx, const CcdParams& P, int* nctas_out) {
    const int p = P.p;
    if (P.ld != ccd_stream_ld(p)) return set_error(ctx, KB_ERR_INVALID, "ccd_stream_run: bad leading dimension");
    if (p > 65535 || (long)P.ld * p > 0x7fffffffL)
        return set_error(ctx, KB_ERR_INVALID, "p = %d too large for the streaming CCD kernel", p);
    // launch shape: 2 CTAs x 384 threads per SM, 4 slots in flight per warp -- the best of the variants
    // measured on B200 at p = 5000 (profiles/r01_ccd_variants.md); KB_CCD_SHAPE=D/F select the alternatives.
    static const char shape = [] {
        const char* e = getenv("KB_CCD_SHAPE");
        return e ? e[0] : 'C';
    }();
    switch (shape) {
        case 'D': return ccd_stream_launch<512, 2, 2, false>(ctx, P, nctas_out);
        case 'F': return ccd_stream_launch<256, 2, 4, false>(ctx, P, nctas_out);
        default: return ccd_stream_launch<384, 2, 4, false>(ctx, P, nctas_out);
    }
}

}  // namespace kb
