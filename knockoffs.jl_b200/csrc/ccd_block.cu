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
//     M ← M + U·diag(c)·Uᵀ,   U = P₀·[t_0 … t_{b-1}],
// i.e. one p x b x b GEMM and one SYRK-shaped p x p x b GEMM on the FP64 tensor cores per block: the
// triangle is read and written ONCE per 64 coordinates instead of once per coordinate.
// Per block: gather (panel + partial Grams, all SMs) → sequential (one CTA, b line searches, registers +
// shared memory only) → U = P₀·T (DMMA GEMM) → rank-b update of the lower tiles (DMMA GEMM).
// The coordinate order, the line search (ccd_device.cuh) and every skip / clip / error decision are the
// reference's; only the order in which the mathematically identical updates reach memory differs.
#include "kb_common.cuh"
#include "ccd.cuh"
#include "ccd_device.cuh"
#include "linalg.cuh"
#include <cstdlib>
#include <vector>

namespace kb {

constexpr int BW = CCD_BLOCK_WIDTH;    // coordinates per block (= slot rows, 64)
constexpr int SLAB = 128;              // rows per gather CTA
constexpr int GPITCH = SLAB + 1;       // shared-memory pitch of one panel column
constexpr int BLK_THREADS = 256;
static_assert(BW == 64 && SLOT_ROWS == 64, "ownership maps below assume 64-wide blocks");

// Thread t of the 256 owns the 4 x 4 entries (tr + 16a, tc + 16b), tr = t & 15, tc = t >> 4, of every 64 x 64
// block matrix (D, G, T) -- in the gather kernel's partial Grams and in the sequential kernel alike, so partial
// Grams are exchanged as [slab][a*4+b][t] (coalesced).

// Panel P₀ = M[:, j0 .. j0+63] (full height, from the lower triangle + symmetric diagonal blocks), written to
// Pbuf (ld x 64, rows >= p zero), and -- WITH_GRAM -- this slab's partial Gram.
template <bool WITH_GRAM>
__global__ void __launch_bounds__(BLK_THREADS) ccd_block_gather_kernel(const double* __restrict__ M, long ld, int p, int j0,
                                                                      double* __restrict__ Pbuf, double* __restrict__ Gpart) {
    extern __shared__ double gsm[];          // [64][GPITCH]: column c of the slab at gsm + c * GPITCH
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * SLAB;
    const int jb = j0 / BW;
#pragma unroll
    for (int h = 0; h < SLAB / BW; ++h) {
        const int rbase = r0 + h * BW;
        const int rb = rbase / BW;
        if (rbase >= ld) {
            for (int e = tid; e < BW * BW; e += BLK_THREADS) gsm[(e >> 6) * GPITCH + h * BW + (e & 63)] = 0.0;
        } else if (rb >= jb) {
            // stored directly: M(r, j0 + c), r fastest (the diagonal block is stored symmetric)
            for (int e = tid; e < BW * BW; e += BLK_THREADS) {
                const int r = e & 63, c = e >> 6;
                gsm[c * GPITCH + h * BW + r] = (j0 + c < p) ? M[(long)(rbase + r) + (long)(j0 + c) * ld] : 0.0;
            }
        } else {
            // rows above the block: M(r, j0 + c) = M(j0 + c, r), c fastest
            for (int e = tid; e < BW * BW; e += BLK_THREADS) {
                const int c = e & 63, r = e >> 6;
                gsm[c * GPITCH + h * BW + r] = (j0 + c < p) ? M[(long)(j0 + c) + (long)(rbase + r) * ld] : 0.0;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < SLAB * BW; e += BLK_THREADS) {
        const int r = e & (SLAB - 1), c = e >> 7;
        if (r0 + r < ld) Pbuf[(long)(r0 + r) + (long)c * ld] = gsm[c * GPITCH + r];
    }
    if (WITH_GRAM) {
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
}

// One CTA walks the block's coordinates in order.  TT (64 x 128, ld 64) receives [t_0 … t_63 | c_0 t_0 … c_63 t_63].
// sweep bookkeeping (counters[3] running max|δ| of the sweep, counters[1] accepted updates, *status) lives in
// device memory across the sweep's blocks; `first` resets it, `last` closes the sweep (trace, λ decay, sweeps_out).
template <bool MVR>
__global__ void __launch_bounds__(BLK_THREADS, 1)
ccd_block_seq_kernel(const CcdParams P, int j0, int bw, const double* __restrict__ Pbuf, const double* __restrict__ Gpart,
                     int nslab, double* __restrict__ TT, int first, int last) {
    __shared__ double wbuf[2][BW], gbuf[2][BW], tbuf[2][BW];
    __shared__ double s_s[BW], s_a[BW], s_k[BW];
    const int tid = threadIdx.x;
    const int tr = tid & 15, tc = tid >> 4;
    const long ld = P.ld;
    double D[4][4], G[4][4], T[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = tr + 16 * a, c = tc + 16 * b;
            D[a][b] = (r < bw && c < bw) ? Pbuf[(long)(j0 + r) + (long)c * ld] : 0.0;
            T[a][b] = (r == c) ? 1.0 : 0.0;
            G[a][b] = 0.0;
        }
    if (MVR) {
        for (int sl = 0; sl < nslab; ++sl) {
            const double* gp = Gpart + (size_t)sl * (BW * BW);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) G[a][b] += gp[(a * 4 + b) * BLK_THREADS + tid];
        }
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
    const int method = P.method;
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
            const double ss = MVR ? gbuf[pb][j] : 0.0;
            const double sj = s_s[j], ajj = s_a[j];
            double delta, s_new, cfac;
            bool skip;
            ccd_line_search(method, m, lam, lambda_min, sj, ajj, s_k[j], ss, Mjj, delta, s_new, skip, cfac, status);
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
                    wr[a] = wbuf[pb][tr + 16 * a];
                    wc[a] = wbuf[pb][tc + 16 * a];
                    tv[a] = cfac * tbuf[pb][tr + 16 * a];
                    if (MVR) {
                        // G += α wᵀ + β gᵀ,  α = c g + c² ss w,  β = c w
                        al[a] = cfac * gbuf[pb][tr + 16 * a] + cfac * cfac * ss * wr[a];
                        gc[a] = gbuf[pb][tc + 16 * a];
                    }
                    wr[a] *= cfac;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        D[a][b] = fma(wr[a], wc[b], D[a][b]);
                        T[a][b] = fma(tv[a], wc[b], T[a][b]);
                        if (MVR) G[a][b] = fma(al[a], wc[b], fma(wr[a], gc[b], G[a][b]));
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
                P.counters[2] = (method == KB_METHOD_SDP_CCD) ? lam * P.sdp_mu : lam;   // utilities.jl:339
            }
            *P.sweeps_out = 1;
        }
    }
}

size_t ccd_block_work_doubles(int p) {
    const long ld = ccd_stream_ld(p);
    const long nslab = (ld + SLAB - 1) / SLAB;
    return (size_t)ld * BW            // Pbuf
           + (size_t)ld * 2 * BW      // [U | U diag(c)]
           + (size_t)nslab * BW * BW  // partial Grams
           + 2 * BW * BW + 64;        // TT
}

// ONE sweep over all coordinates (P.nsweeps is ignored; P.sweep0 / P.sdp_lambda describe this sweep).
int ccd_block_run(kb_ctx* ctx, const CcdParams& P, double* work) {
    const int p = P.p;
    const long ld = P.ld;
    if (ld != ccd_stream_ld(p)) return set_error(ctx, KB_ERR_INVALID, "ccd_block_run: bad leading dimension");
    const int nslab = (int)((ld + SLAB - 1) / SLAB);
    double* Pbuf = work;
    double* UU = Pbuf + (size_t)ld * BW;
    double* Gpart = UU + (size_t)ld * 2 * BW;
    double* TT = Gpart + (size_t)nslab * BW * BW;
    const bool mvr = (P.method == KB_METHOD_MVR);
    const size_t gsmem = (size_t)BW * GPITCH * sizeof(double);
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)ccd_block_gather_kernel<true>, gsmem));
    KB_CUDA(ctx, raise_max_dynamic_smem((const void*)ccd_block_gather_kernel<false>, gsmem));
    const int nblk = (p + BW - 1) / BW;
    // KB_BLOCK_PHASES=1: CUDA events around every launch of the sweep, per-phase totals on stderr (diagnostics)
    static const bool phases = getenv("KB_BLOCK_PHASES") != nullptr;
    std::vector<cudaEvent_t> ev;
    if (phases) {
        ev.resize((size_t)nblk * 5);
        for (auto& e : ev) cudaEventCreate(&e);
    }
#define KB_EV(i) do { if (phases) cudaEventRecord(ev[(size_t)blk * 5 + (i)], ctx->stream); } while (0)
    for (int blk = 0; blk < nblk; ++blk) {
        const int j0 = blk * BW;
        const int bw = (p - j0 < BW) ? p - j0 : BW;
        const int first = (blk == 0), last = (blk == nblk - 1);
        KB_EV(0);
        if (mvr) ccd_block_gather_kernel<true><<<nslab, BLK_THREADS, gsmem, ctx->stream>>>(P.M, ld, p, j0, Pbuf, Gpart);
        else ccd_block_gather_kernel<false><<<nslab, BLK_THREADS, gsmem, ctx->stream>>>(P.M, ld, p, j0, Pbuf, Gpart);
        KB_EV(1);
        if (mvr) ccd_block_seq_kernel<true><<<1, BLK_THREADS, 0, ctx->stream>>>(P, j0, bw, Pbuf, Gpart, nslab, TT, first, last);
        else ccd_block_seq_kernel<false><<<1, BLK_THREADS, 0, ctx->stream>>>(P, j0, bw, Pbuf, Gpart, nslab, TT, first, last);
        KB_EV(2);
        ctx->launches += 2;
        KB_CUDA(ctx, cudaGetLastError());
        // [U | U diag(c)] = P₀ · TT
        KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, p, 2 * BW, BW, 1.0, Pbuf, ld, TT, BW, 0.0, UU, ld));
        KB_EV(3);
        // M (lower tiles) += (U diag(c)) · Uᵀ
        KB_TRY(gemm1(ctx, A_MCONTIG, B_NCONTIG, p, p, BW, 1.0, UU + (size_t)ld * BW, ld, UU, ld, 1.0, P.M, ld, 0, 1, 0));
        KB_EV(4);
    }
#undef KB_EV
    if (phases) {
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        double t[4] = {0, 0, 0, 0};
        for (int blk = 0; blk < nblk; ++blk)
            for (int i = 0; i < 4; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)blk * 5 + i], ev[(size_t)blk * 5 + i + 1]);
                t[i] += ms;
            }
        float tot = 0.f;
        cudaEventElapsedTime(&tot, ev[0], ev[(size_t)nblk * 5 - 1]);
        fprintf(stderr, "[ccd_block sweep: %d blocks, %.2f ms] gather=%.2f seq=%.2f U=P*T=%.2f update=%.2f ms\n", nblk, tot, t[0], t[1],
                t[2], t[3]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return KB_OK;
}

}  // namespace kb
