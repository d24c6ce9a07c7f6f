// Blocked (delayed-update) coordinate descent for solve_MVR / solve_max_entropy / solve_sdp_ccd
// (src/utilities.jl:124-175, :221-280, :291-343) on the resident inverse M = A⁻¹, A = κΣ − S (+ λmin I).
//
// ccd_stream.cu applies every coordinate's Sherman–Morrison update  M ← M + c·m_j m_jᵀ  to the whole triangle at
// once (8p² B of HBM traffic per coordinate).  The same sequence of rank-1 updates can be applied lazily: inside
// a block J of b = 64 consecutive coordinates the line searches only need
//     M_jj                (the current diagonal entry)           -> the b x b block D = M[J, J], kept current,
//     ‖M[:, j]‖²  (MVR)    (the current column's squared norm)    -> the b x b Gram  G = M[:, J]ᵀ M[:, J], kept current,
// and both obey O(b²) recurrences under a rank-1 update with u = M[:, j], w = u[J] = D[:, j], g = G[:, j]:
//     D ← D + c·w wᵀ,     G ← G + c·(g wᵀ + w gᵀ) + c²·G_jj·w wᵀ.
// The current panel is M[:, J] = P₀·T with P₀ the panel at block entry and T (b x b) following T ← T + c·T[:, j]·wᵀ,
// so u_j = P₀·T[:, j] (snapshot at step j) and the whole block amounts to the rank-b update
//     M ← M + U·diag(c)·Uᵀ,   U = P₀·[t_0 … t_{b-1}].
// Two levels: an OUTER block of up to 4 inner blocks (256 coordinates) keeps a gathered full-height copy PO of its
// columns; after each inner block only PO's later columns are brought up to date (p x ≤192 x 64 GEMM), and the
// triangle itself is read and written ONCE per outer block by a SYRK-shaped p x p x 256 GEMM on the FP64 tensor
// cores -- instead of once per coordinate.
// Per inner block: [Gram partials (MVR)] → sequential kernel (one CTA, 64 line searches, registers + shared memory
// only) → [U | U diag(c)] = P₀·[T | T diag(c)] (DMMA GEMM) → update of PO's later columns (DMMA GEMM).
// The coordinate order, the line search (ccd_device.cuh) and every skip / clip / error decision are the
// reference's; only the order in which the mathematically identical updates reach memory differs.
#include "kb_common.cuh"
#include "ccd.cuh"
#include "ccd_device.cuh"
#include "linalg.cuh"
#include <cstdlib>
#include <vector>
#include <nvtx3/nvToolsExt.h>     // header-only; ranges are emitted only with KB_NVTX=1 (ncu --nvtx --nvtx-include "ccd_big/")

namespace kb {

constexpr int BW = CCD_BLOCK_WIDTH;    // coordinates per inner block (= slot rows, 64)
constexpr int MAX_NIN = 8;             // inner blocks per outer block: at most 8 (512 coordinates)
constexpr int SLAB = 128;              // rows per gather / Gram CTA
constexpr int GPITCH = SLAB + 1;       // shared-memory pitch of one panel column
constexpr int BLK_THREADS = 256;
#ifndef KB_CHAIN_FAST_MATH
#define KB_CHAIN_FAST_MATH 1
#endif
constexpr bool CHAIN_FAST_MATH = KB_CHAIN_FAST_MATH != 0;   // MUFU-seeded reciprocal / sqrt on the sequential kernel's chain (ccd_device.cuh)
constexpr int GRAM_SLOTS = 5;                    // double2 rows per thread and pass of the direct column-norm recompute
constexpr double GRAM_RECOMPUTE_RATIO = 2e-6;    // recompute G_jj when it is below this fraction of its error scale
static_assert(BW == 64 && SLOT_ROWS == 64, "ownership maps below assume 64-wide blocks");

// Thread t of the 256 owns the 4 x 4 entries (tr + 16a, tc + 16b), tr = t & 15, tc = t >> 4, of every 64 x 64
// block matrix (D, G, T) -- in the Gram kernel's partials and in the sequential kernel alike, so partial
// Grams are exchanged as [slab][a*4+b][t] (coalesced).

// PO[:, 64 cb .. 64 cb + 63] = M[:, j0 .. j0 + 63], j0 = o0 + 64 cb (full height, from the lower triangle + the
// symmetric diagonal blocks; rows >= p and columns >= p are zero).  grid = (row slabs, inner blocks).
__global__ void __launch_bounds__(BLK_THREADS) ccd_block_gather_kernel(const double* __restrict__ M, long ld, int p, int o0,
                                                                      double* __restrict__ PO) {
    __shared__ double tile[BW][BW + 1];
    const int tid = threadIdx.x;
    const int j0 = o0 + (int)blockIdx.y * BW;
    const int jb = j0 / BW;
    double* Pk = PO + (size_t)blockIdx.y * BW * ld;
#pragma unroll
    for (int h = 0; h < SLAB / BW; ++h) {
        const int rbase = (int)blockIdx.x * SLAB + h * BW;
        if (rbase >= ld) break;
        const int rb = rbase / BW;
        if (rb >= jb) {
            // stored directly: M(r, j0 + c), r fastest (the diagonal block is stored symmetric)
            for (int e = tid; e < BW * BW; e += BLK_THREADS) {
                const int r = e & 63, c = e >> 6;
                Pk[(long)(rbase + r) + (long)c * ld] = (j0 + c < p) ? M[(long)(rbase + r) + (long)(j0 + c) * ld] : 0.0;
            }
        } else {
            // rows above the block: M(r, j0 + c) = M(j0 + c, r): read c fastest, transpose through shared memory
            __syncthreads();
            for (int e = tid; e < BW * BW; e += BLK_THREADS) {
                const int c = e & 63, r = e >> 6;
                tile[r][c] = (j0 + c < p) ? M[(long)(j0 + c) + (long)(rbase + r) * ld] : 0.0;
            }
            __syncthreads();
            for (int e = tid; e < BW * BW; e += BLK_THREADS) {
                const int r = e & 63, c = e >> 6;
                Pk[(long)(rbase + r) + (long)c * ld] = tile[r][c];
            }
        }
    }
}

// partial Gram of one 128-row slab of the current panel Pk (ld x 64)
__global__ void __launch_bounds__(BLK_THREADS) ccd_block_gram_kernel(const double* __restrict__ Pk, long ld,
                                                                    double* __restrict__ Gpart) {
    extern __shared__ double gsm[];          // [64][GPITCH]: column c of the slab at gsm + c * GPITCH
    const int tid = threadIdx.x;
    const int r0 = (int)blockIdx.x * SLAB;
    for (int e = tid; e < SLAB * BW; e += BLK_THREADS) {
        const int r = e & (SLAB - 1), c = e >> 7;
        gsm[c * GPITCH + r] = (r0 + r < ld) ? Pk[(long)(r0 + r) + (long)c * ld] : 0.0;
    }
    __syncthreads();
    const int tr = tid & 15, tc = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const double* pa = gsm + tr * GPITCH;
    const double* pb = gsm + tc * GPITCH;
#pragma unroll 4
    for (int r = 0; r < SLAB; ++r) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            av[a] = pa[a * 16 * GPITCH + r];
            bv[a] = pb[a * 16 * GPITCH + r];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    double* out = Gpart + (size_t)blockIdx.x * (BW * BW);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[(a * 4 + b) * BLK_THREADS + tid] = acc[a][b];
}

// One CTA walks the block's coordinates in order.  TT (64 x 128, ld 64) receives [t_0 … t_63 | c_0 t_0 … c_63 t_63].
// sweep bookkeeping (counters[3] running max|δ| of the sweep, counters[1] accepted updates, *status) lives in
// device memory across the sweep's blocks; `first` resets it, `last` closes the sweep (trace, λ decay, sweeps_out).
// After step j only D / G entries with both indices > j and T entries (row <= j, column > j) are read again, so the
// register tiles are updated from the current 16-column phase bq onwards only (compile-time bounds).
template <int METHOD>
__global__ void __launch_bounds__(BLK_THREADS, 1)
ccd_block_seq_kernel(const CcdParams P, int j0, int bw, const double* __restrict__ Pk, const double* __restrict__ Gpart,
                     int nslab, double* __restrict__ TT, int first, int last) {
    constexpr bool MVR = (METHOD == KB_METHOD_MVR);
    __shared__ double wbuf[2][BW], gbuf[2][BW], tbuf[2][BW];
    __shared__ double s_s[BW], s_a[BW], s_k[BW];
    __shared__ double gsc[BW];      // MVR: magnitude of everything summed into G_jj so far (its rounding-error scale)
    __shared__ double red[BLK_THREADS / 32];
    const int tid = threadIdx.x;
    const int tr = tid & 15, tc = tid >> 4;
    const long ld = P.ld;
    double D[4][4], G[4][4], T[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = tr + 16 * a, c = tc + 16 * b;
            D[a][b] = (r < bw && c < bw) ? Pk[(long)(j0 + r) + (long)c * ld] : 0.0;
            T[a][b] = (r == c) ? 1.0 : 0.0;
            G[a][b] = 0.0;
        }
    if (MVR) {
        // partial Grams of the 128-row slabs, summed in slab order (deterministic).  Four slabs per round: 64 independent
        // loads in flight per thread -- one L2 round trip per four slabs instead of one per slab (ncu r01b: this loop was
        // the kernel's long-scoreboard stall, ~18 % of its 67 us at p = 5000).
        int sl = 0;
        for (; sl + 4 <= nslab; sl += 4) {
            double t[4][16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* gp = Gpart + (size_t)(sl + u) * (BW * BW);
#pragma unroll
                for (int e = 0; e < 16; ++e) t[u][e] = gp[e * BLK_THREADS + tid];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) G[a][b] += t[u][a * 4 + b];
        }
        for (; sl < nslab; ++sl) {
            const double* gp = Gpart + (size_t)sl * (BW * BW);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) G[a][b] += gp[(a * 4 + b) * BLK_THREADS + tid];
        }
    }
    if (MVR && tr == tc) {
#pragma unroll
        for (int a = 0; a < 4; ++a) gsc[tr + 16 * a] = fabs(G[a][a]);
    }
    if (tid < BW) {
        const bool in = tid < bw;
        s_s[tid] = in ? P.s[j0 + tid] : 0.0;
        s_a[tid] = in ? P.adiag[j0 + tid] : 1.0;
        s_k[tid] = (in && !MVR) ? P.ksig_diag[j0 + tid] : 0.0;
    }
    int status = *P.status;
    double max_delta = first ? 0.0 : P.counters[3];
    double nupd = 0.0;
    const double m = P.m, lam = P.sdp_lambda, lambda_min = P.lambda_min;
    __syncthreads();

#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
        for (int jj = 0; jj < 16; ++jj) {
            const int j = bq * 16 + jj;
            if (j >= bw) break;                      // uniform
            const int pb = j & 1;
            if (tc == jj) {                          // owners of column j publish it
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    wbuf[pb][tr + 16 * a] = D[a][bq];
                    if (MVR) gbuf[pb][tr + 16 * a] = G[a][bq];
                    tbuf[pb][tr + 16 * a] = T[a][bq];
                }
            }
            __syncthreads();
            const double Mjj = wbuf[pb][j];
            double ss = MVR ? gbuf[pb][j] : 0.0;
            if (MVR && ss < GRAM_RECOMPUTE_RATIO * gsc[j]) {
                // The recurrence for G_jj = ‖M[:, j]‖² has cancelled more than ~6 digits (first sweep from the equi start:
                // A is within λmin of singular, the columns of M shrink from ~1/λmin to O(1) as a near-null group is
                // eliminated): recompute it directly from the data, ‖P₀ t_j‖², where only ~6 digits cancel.  The branch is
                // uniform (every thread reads the same shared values); ≈ one pass of this CTA over the p x (j+1) panel.
                double acc = 0.0;
                const double* tj = tbuf[pb];
                for (long base = 2 * tid; base < ld; base += 2L * BLK_THREADS * GRAM_SLOTS) {
                    double2 u[GRAM_SLOTS];
#pragma unroll
                    for (int i = 0; i < GRAM_SLOTS; ++i) u[i] = make_double2(0.0, 0.0);
                    const double* col = Pk + base;
#pragma unroll 2
                    for (int k = 0; k <= j; ++k, col += ld) {
                        const double t = tj[k];
#pragma unroll
                        for (int i = 0; i < GRAM_SLOTS; ++i) {
                            if (base + 2L * BLK_THREADS * i < ld) {
                                const double2 v = *reinterpret_cast<const double2*>(col + 2 * BLK_THREADS * i);
                                u[i].x = fma(v.x, t, u[i].x);
                                u[i].y = fma(v.y, t, u[i].y);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < GRAM_SLOTS; ++i) acc = fma(u[i].x, u[i].x, fma(u[i].y, u[i].y, acc));
                }
                ss = block_sum<BLK_THREADS / 32>(acc, red, tid);
            }
            const double sj = s_s[j], ajj = s_a[j];
            double delta, s_new, cfac;
            bool skip;
            ccd_line_search<CHAIN_FAST_MATH>(METHOD, m, lam, lambda_min, sj, ajj, s_k[j], ss, Mjj, delta, s_new, skip, cfac, status);
            const bool upd = !skip && status == 0;
            if (upd) {
                if (fabs(delta) > max_delta) max_delta = fabs(delta);
                nupd += 1.0;
                if (tid == 0) {
                    P.s[j0 + j] = s_new;
                    P.adiag[j0 + j] = ajj - delta;
                }
            }
            if (tc == jj) {                          // snapshot t_j and c_j t_j (rows > j of t_j are zero)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double tv = T[a][bq];
                    TT[(tr + 16 * a) + (long)j * BW] = tv;
                    TT[(tr + 16 * a) + (long)(BW + j) * BW] = upd ? cfac * tv : 0.0;
                }
            }
            if (upd) {
                double wr[4], wc[4], al[4], tv[4], gc[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a >= bq) {
                        wc[a] = wbuf[pb][tc + 16 * a];
                        const double w = wbuf[pb][tr + 16 * a];
                        if (MVR) {
                            // G += α wᵀ + β gᵀ,  α = c g + c² ss w,  β = c w
                            al[a] = cfac * gbuf[pb][tr + 16 * a] + cfac * cfac * ss * w;
                            gc[a] = gbuf[pb][tc + 16 * a];
                        }
                        wr[a] = cfac * w;
                        if (MVR && tr == tc && tr + 16 * a > j)        // |terms| added to G_ii below (i > j only: gsc[j] is being read)
                            gsc[tr + 16 * a] += fabs(cfac) * 2.0 * fabs(gbuf[pb][tr + 16 * a] * w) + cfac * cfac * ss * w * w;
                    }
                    if (a <= bq) tv[a] = cfac * tbuf[pb][tr + 16 * a];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (b < bq) continue;
                        if (a >= bq) {
                            D[a][b] = fma(wr[a], wc[b], D[a][b]);
                            if (MVR) G[a][b] = fma(al[a], wc[b], fma(wr[a], gc[b], G[a][b]));
                        }
                        if (a <= bq) T[a][b] = fma(tv[a], wc[b], T[a][b]);
                    }
            }
        }
    }
    // columns j >= bw of a ragged last block carry no update
    for (int e = tid; e < BW * BW; e += BLK_THREADS) {
        const int j = e >> 6;
        if (j >= bw) {
            TT[e] = 0.0;
            TT[e + BW * BW] = 0.0;
        }
    }
    if (tid == 0) {
        P.counters[3] = max_delta;
        P.counters[1] += nupd;
        *P.status = status;
        if (last) {
            if (status == 0) {
                if (P.sweep0 < KB_MAX_TRACE) P.trace[P.sweep0] = max_delta;
                P.counters[2] = (METHOD == KB_METHOD_SDP_CCD) ? lam * P.sdp_mu : lam;   // utilities.jl:339
            }
            *P.sweeps_out = 1;
        }
    }
}

static int outer_blocks() {
    static const int v = [] {
        const char* e = getenv("KB_CCD_OUTER");      // inner blocks per outer block (tuning), default 4
        int x = e ? atoi(e) : 4;
        return x < 1 ? 1 : (x > MAX_NIN ? MAX_NIN : x);
    }();
    return v;
}

size_t ccd_block_work_doubles(int p) {
    const long ld = ccd_stream_ld(p);
    const long nslab = (ld + SLAB - 1) / SLAB;
    const long wo = (long)BW * outer_blocks();
    return (size_t)ld * wo * 5              // PO, 2 x (U_all, Uc_all): ping-pong so a rank-256 pass can run beside the next block
           + (size_t)nslab * BW * BW        // partial Grams
           + 2 * BW * BW + 64;              // TT
}

// ONE sweep over all coordinates (P.nsweeps is ignored; P.sweep0 / P.sdp_lambda describe this sweep).
int ccd_block_run(kb_ctx* ctx, const CcdParams& P, double* work) {
    const int p = P.p;
    const long ld = P.ld;
    if (ld != ccd_stream_ld(p)) return set_error(ctx, KB_ERR_INVALID, "ccd_block_run: bad leading dimension");
    const int nslab = (int)((ld + SLAB - 1) / SLAB);
    const int NIN = outer_blocks();
    const long wo = (long)BW * NIN;
    double* PO = work;
    double* Ubuf[2] = {PO + (size_t)ld * wo, PO + 3 * (size_t)ld * wo};
    double* Ucbuf[2] = {PO + 2 * (size_t)ld * wo, PO + 4 * (size_t)ld * wo};
    double* Gpart = PO + 5 * (size_t)ld * wo;
    double* TT = Gpart + (size_t)nslab * BW * BW;
    const bool mvr = (P.method == KB_METHOD_MVR);
    const size_t gsmem = (size_t)BW * GPITCH * sizeof(double);
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)ccd_block_gram_kernel, gsmem));
    const int nblk = (p + BW - 1) / BW;
    // KB_BLOCK_PHASES=1: CUDA events around every phase of the sweep, per-phase totals on stderr (diagnostics; serial order)
    static const bool phases = getenv("KB_BLOCK_PHASES") != nullptr;
    // Overlap (default): the rank-256 pass over the triangle of outer block ob runs on the low-priority auxiliary stream
    // WHILE the main stream walks outer block ob + 1, whose panel is formed from the not-yet-updated triangle plus the
    // same rank-256 term restricted to its 256 columns (a p x 256 x 256 GEMM).  The single-CTA sequential kernels then
    // no longer leave the other SMs idle.  KB_CCD_OVERLAP=0 restores the serial order.
    static const bool overlap_env = [] { const char* e = getenv("KB_CCD_OVERLAP"); return !(e && e[0] == '0'); }();
    const bool overlap = overlap_env && !phases;
    enum { PH_GATHER, PH_GRAM, PH_SEQ, PH_PT, PH_INNER, PH_BIG, PH_N };
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_phase;
    auto mark = [&](int phase) {
        if (!phases) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        ev.push_back(e);
        ev_phase.push_back(phase);
    };
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream_aux;
    cudaEvent_t ev_u = ctx->ccd_ev[0], ev_g = ctx->ccd_ev[1], ev_big = ctx->ccd_ev[2];
    // rows p..ld-1 of U_all / Uc_all are never written by the GEMMs (M = p) but are read as operand rows of the panel
    // update: keep them zero
    KB_CUDA(ctx, cudaMemsetAsync(Ubuf[0], 0, sizeof(double) * 4 * (size_t)ld * wo, s0));
    mark(-1);
    bool big_pending = false;
    const int nouter = (nblk + NIN - 1) / NIN;
    for (int ob = 0; ob < nouter; ++ob) {
        const int o0 = ob * NIN * BW;
        const int nin = (nblk - ob * NIN < NIN) ? nblk - ob * NIN : NIN;
        double* Uall = Ubuf[ob & 1];
        double* Ucall = Ucbuf[ob & 1];
        if (ob == 0 || !overlap) {
            ccd_block_gather_kernel<<<dim3(nslab, nin), BLK_THREADS, 0, s0>>>(P.M, ld, p, o0, PO);
            ctx->launches++;
            mark(PH_GATHER);
        }   // else: PO was formed at the end of the previous iteration
        for (int k = 0; k < nin; ++k) {
            const int blk = ob * NIN + k;
            const int j0 = blk * BW;
            const int bw = (p - j0 < BW) ? p - j0 : BW;
            const int first = (blk == 0), last = (blk == nblk - 1);
            double* Pk = PO + (size_t)k * BW * ld;
            if (mvr) {
                ccd_block_gram_kernel<<<nslab, BLK_THREADS, gsmem, s0>>>(Pk, ld, Gpart);
                ctx->launches++;
                mark(PH_GRAM);
            }
            if (P.method == KB_METHOD_MVR)
                ccd_block_seq_kernel<KB_METHOD_MVR><<<1, BLK_THREADS, 0, s0>>>(P, j0, bw, Pk, Gpart, nslab, TT, first, last);
            else if (P.method == KB_METHOD_MAXENT)
                ccd_block_seq_kernel<KB_METHOD_MAXENT><<<1, BLK_THREADS, 0, s0>>>(P, j0, bw, Pk, Gpart, nslab, TT, first, last);
            else
                ccd_block_seq_kernel<KB_METHOD_SDP_CCD><<<1, BLK_THREADS, 0, s0>>>(P, j0, bw, Pk, Gpart, nslab, TT, first, last);
            ctx->launches++;
            KB_CUDA(ctx, cudaGetLastError());
            mark(PH_SEQ);
            // [U_k | U_k diag(c)] = P₀ · TT   ->  columns 64k.. of U_all and Uc_all
            {
                GemmParams G{};
                G.M = p; G.N = 2 * BW; G.nseg = 1;
                G.seg[0] = GemmSeg{Pk, ld, TT, BW, BW, 0};
                G.C = Uall + (size_t)k * BW * ld; G.ldc = ld; G.Cin = G.C; G.ldcin = ld;
                G.C2 = Ucall + (size_t)k * BW * ld; G.nsplit = BW;
                G.alpha = 1.0; G.beta = 0.0;
                KB_TRY(gemm(ctx, A_MCONTIG, B_KCONTIG, G));
            }
            mark(PH_PT);
            // PO's later columns += (U_k diag(c)) · U_k[rows of those columns, :]ᵀ
            if (k + 1 < nin) {
                const int nlater = (nin - 1 - k) * BW;
                KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, p, nlater, BW, 1.0, Ucall + (size_t)k * BW * ld, ld,
                             Uall + (size_t)k * BW * ld + (size_t)(j0 + BW), ld, 1.0, PO + (size_t)(k + 1) * BW * ld, ld));
                mark(PH_INNER);
            }
        }
        if (!overlap) {
            // M (lower tiles) += Uc_all · U_allᵀ   (K = 64 nin)
            static const bool nvtx_on = getenv("KB_NVTX") != nullptr;
            if (nvtx_on) nvtxRangePushA("ccd_big");
            const int rc_big = gemm1(ctx, A_MCONTIG, B_NCONTIG, p, p, nin * BW, 1.0, Ucall, ld, Uall, ld, 1.0, P.M, ld, 0, 1, 0);
            if (nvtx_on) nvtxRangePop();
            KB_TRY(rc_big);
            mark(PH_BIG);
            continue;
        }
        if (ob + 1 < nouter) {
            // next panel = (triangle with the passes up to ob − 1 applied) + this block's rank-256 term on its columns
            const int o1 = (ob + 1) * NIN * BW;
            const int nin1 = (nblk - (ob + 1) * NIN < NIN) ? nblk - (ob + 1) * NIN : NIN;
            if (big_pending) KB_CUDA(ctx, cudaStreamWaitEvent(s0, ev_big, 0));
            ccd_block_gather_kernel<<<dim3(nslab, nin1), BLK_THREADS, 0, s0>>>(P.M, ld, p, o1, PO);
            ctx->launches++;
            KB_CUDA(ctx, cudaEventRecord(ev_g, s0));
            KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, p, nin1 * BW, nin * BW, 1.0, Ucall, ld, Uall + (size_t)o1, ld, 1.0, PO, ld));
            KB_CUDA(ctx, cudaStreamWaitEvent(s1, ev_g, 0));
        } else {
            KB_CUDA(ctx, cudaEventRecord(ev_u, s0));
            KB_CUDA(ctx, cudaStreamWaitEvent(s1, ev_u, 0));
        }
        ctx->stream = s1;
        const int rc = gemm1(ctx, A_MCONTIG, B_NCONTIG, p, p, nin * BW, 1.0, Ucall, ld, Uall, ld, 1.0, P.M, ld, 0, 1, 0);
        ctx->stream = s0;
        KB_TRY(rc);
        KB_CUDA(ctx, cudaEventRecord(ev_big, s1));
        big_pending = true;
    }
    if (big_pending) KB_CUDA(ctx, cudaStreamWaitEvent(s0, ev_big, 0));
    if (phases) {
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        double t[PH_N] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            t[ev_phase[i]] += ms;
        }
        float tot = 0.f;
        cudaEventElapsedTime(&tot, ev.front(), ev.back());
        fprintf(stderr, "[ccd_block sweep: %d blocks of 64 in groups of %d, %.2f ms] gather=%.2f gram=%.2f seq=%.2f U=P*T=%.2f "
                        "panel=%.2f rank-%d update=%.2f ms\n", nblk, NIN, tot, t[PH_GATHER], t[PH_GRAM], t[PH_SEQ], t[PH_PT],
                t[PH_INNER], NIN * BW, t[PH_BIG]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return KB_OK;
}

}  // namespace kb
