// Representative group knockoffs: choose_group_reps (src/group.jl:1929-2033), cond_indep_corr
// (:205-240), solve_s_graphical_group (:268-303), modelX_gaussian_rep_group_knockoffs (:170-203).
//
// The p-sized algebra runs on the device (inv(Σ), inv(Σ11), the r x (p−r) and (p−r) x (p−r) products on the DMMA
// GEMM); the per-group decisions of choose_group_reps are g x g host work on the group-diagonal blocks of Σ and
// inv(Σ).  The reference forms inv(Σ[RO, RO]) (a (p − |ROc|)-sized matrix, once per candidate step) only to
// evaluate x' inv(Σ[RO,RO]) x for columns j of the group; with Q = inv(inv(Σ)[ROc, ROc]) = Cov(ROc | RO) that
// quadratic form is Σ_jj − Q_jj exactly (Schur complement), so only |ROc| x |ROc| inverses are needed.
#include "rep.cuh"
#include "group.cuh"
#include "linalg.cuh"
#include "pipeline.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace kb {

namespace {

inline bool isapprox(double a, double b) {       // Julia isapprox: |a − b| <= sqrt(eps) max(|a|, |b|)
    return std::fabs(a - b) <= 1.4901161193847656e-08 * std::fmax(std::fabs(a), std::fabs(b));
}

// ---- small dense SPD helpers on the host (column-major n x n) ------------------------------------------
bool chol_lower(std::vector<double>& A, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[j + (size_t)j * n];
        for (int k = 0; k < j; ++k) d -= A[j + (size_t)k * n] * A[j + (size_t)k * n];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        A[j + (size_t)j * n] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = A[i + (size_t)j * n];
            for (int k = 0; k < j; ++k) v -= A[i + (size_t)k * n] * A[j + (size_t)k * n];
            A[i + (size_t)j * n] = v / d;
        }
    }
    return true;
}
// A <- inv(A) for SPD A (full symmetric in, full symmetric out)
bool spd_inverse_small(std::vector<double>& A, int n) {
    if (!chol_lower(A, n)) return false;
    std::vector<double> W((size_t)n * n, 0.0);                    // W = inv(L), lower
    for (int j = 0; j < n; ++j) {
        W[j + (size_t)j * n] = 1.0 / A[j + (size_t)j * n];
        for (int i = j + 1; i < n; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= A[i + (size_t)k * n] * W[k + (size_t)j * n];
            W[i + (size_t)j * n] = v / A[i + (size_t)i * n];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = 0.0;
            for (int k = i; k < n; ++k) v += W[k + (size_t)i * n] * W[k + (size_t)j * n];
            A[i + (size_t)j * n] = A[j + (size_t)i * n] = v;
        }
    return true;
}

struct SymView {                                   // Symmetric(Σ), uplo = :U, on a host matrix
    const double* a;
    long ld;
    double operator()(int i, int j) const { return i <= j ? a[i + (size_t)j * ld] : a[j + (size_t)i * ld]; }
};

std::vector<double> gather_block(const SymView& S, const std::vector<int>& rows, const std::vector<int>& cols) {
    std::vector<double> B(rows.size() * cols.size());
    for (size_t j = 0; j < cols.size(); ++j)
        for (size_t i = 0; i < rows.size(); ++i) B[i + j * rows.size()] = S(rows[i], cols[j]);
    return B;
}

// select_one / select_best_rss_subset (group.jl:2070-2098): indices into the group, most important first
std::vector<int> select_best_rss_subset(std::vector<double> C, int p) {
    std::vector<int> indices, vlist(p);
    for (int i = 0; i < p; ++i) vlist[i] = i;
    const double RSS0 = p;
    int n = p;
    for (int it = 0; it < p && n > 0; ++it) {
        int imax = 0;
        double best = -INFINITY, sumd = 0.0;
        std::vector<double> rs(n);
        for (int c = 0; c < n; ++c) {
            double s2 = 0.0;
            for (int r = 0; r < n; ++r) s2 += C[r + (size_t)c * n] * C[r + (size_t)c * n];
            rs[c] = s2 / C[c + (size_t)c * n];
            sumd += C[c + (size_t)c * n];
            if (rs[c] > best) { best = rs[c]; imax = c; }            // findmax: first maximum
        }
        const double vmin = sumd - rs[imax];
        indices.push_back(vlist[imax]);
        const double R2 = 1.0 - vmin / RSS0;
        // residC = C − C[:, imax] C[:, imax]' / C[imax, imax]; keep the rows/cols whose residual diagonal > 1e-12
        std::vector<double> R((size_t)n * n);
        const double piv = C[imax + (size_t)imax * n];
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r)
                R[r + (size_t)c * n] = C[r + (size_t)c * n] - C[r + (size_t)imax * n] * C[c + (size_t)imax * n] / piv;
        std::vector<int> keep;
        for (int c = 0; c < n; ++c)
            if (R[c + (size_t)c * n] > 1e-12) keep.push_back(c);
        const int n2 = (int)keep.size();
        std::vector<double> C2((size_t)n2 * n2);
        std::vector<int> v2(n2);
        for (int c = 0; c < n2; ++c) {
            v2[c] = vlist[keep[c]];
            for (int r = 0; r < n2; ++r) C2[r + (size_t)c * n2] = R[keep[r] + (size_t)keep[c] * n];
        }
        C.swap(C2);
        vlist.swap(v2);
        n = n2;
        if (R2 > 1.0 - 1e-12) break;                                // group.jl:2094
    }
    return indices;
}

// prioritize_variants (group.jl:2050-2056)
std::vector<int> prioritize_variants(const std::vector<int>& index, const std::vector<int>& first_pos) {
    std::vector<int> out;
    std::vector<char> used(index.size(), 0);
    for (int f : first_pos) { out.push_back(index[f]); used[f] = 1; }
    for (size_t i = 0; i < index.size(); ++i)
        if (!used[i]) out.push_back(index[i]);
    return out;
}

__global__ void gather_sym_kernel(const double* __restrict__ Sig, long lds, int p, const int* __restrict__ perm, int upper_only,
                                  double* __restrict__ out, long ldo) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p), j = (int)(idx / p);
        int a = perm[i], b = perm[j];
        if (upper_only && a > b) { const int t = a; a = b; b = t; }
        out[i + (long)j * ldo] = Sig[a + (long)b * lds];
    }
}
__global__ void sub_kernel(const double* __restrict__ A, long lda, const double* __restrict__ B, long ldb, int rows, int cols,
                           double* __restrict__ C, long ldc) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)rows * cols;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        C[i + (long)j * ldc] = A[i + (long)j * lda] - B[i + (long)j * ldb];
    }
}
// Assemble the p x p matrix [B11 B12; B12' B22] given in (reps, non-reps) order back into the original variable
// order, reading the UPPER triangle in the ORIGINAL order (Symmetric(·), uplo = :U) and zeroing |x| < thresh.
__global__ void scatter_blocks_kernel(const double* __restrict__ B11, long ld11, const double* __restrict__ B12, long ld12,
                                      const double* __restrict__ B22, long ld22, int p, int r, const int* __restrict__ iperm,
                                      double thresh, double* __restrict__ out, long ldo) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p * p;
    for (; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int a = (int)(idx % p), b = (int)(idx / p);
        if (a > b) { const int t = a; a = b; b = t; }
        const int ia = iperm[a], ib = iperm[b];
        double v;
        if (ia < r && ib < r) v = B11[ia + (long)ib * ld11];
        else if (ia < r) v = B12[ia + (long)(ib - r) * ld12];
        else if (ib < r) v = B12[ib + (long)(ia - r) * ld12];
        else v = B22[(ia - r) + (long)(ib - r) * ld22];
        if (fabs(v) < thresh) v = 0.0;
        out[idx % p + (idx / p) * ldo] = v;
    }
}

int ewg(kb_ctx* ctx, long total) {
    long g = (total + 255) / 256, cap = (long)ctx->num_sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct Perm {
    std::vector<int> perm, iperm, non;              // perm = [reps; non-reps], iperm[perm[i]] = i
};
int make_perm(kb_ctx* ctx, int p, const std::vector<int>& reps, Perm* P) {
    std::vector<char> isrep(p, 0);
    for (int r : reps) {
        if (r < 0 || r >= p || isrep[r]) return set_error(ctx, KB_ERR_INVALID, "group_reps must be distinct indices in 1..p");
        isrep[r] = 1;
    }
    P->perm = reps;
    for (int i = 0; i < p; ++i)
        if (!isrep[i]) { P->perm.push_back(i); P->non.push_back(i); }
    P->iperm.assign(p, 0);
    for (int i = 0; i < p; ++i) P->iperm[P->perm[i]] = i;
    return KB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// choose_group_reps (group.jl:1929-2033).  Sigma: host p x p, upper triangle trusted, must be a correlation
// matrix.  groups: p ids.  prioritize: 0-based indices (may be empty, `has_prior` distinguishes `nothing`).
// reps_out: sorted 0-based indices.
// ------------------------------------------------------------------------------------------------------
int choose_group_reps_host(kb_ctx* ctx, const double* Sigma, long lds, int p, const int64_t* groups, double threshold,
                           bool has_prior, const std::vector<int>& prioritize, std::vector<int>* reps_out) {
    SymView S{Sigma, lds};
    static const bool trace = getenv("KB_REP_TRACE") != nullptr;     // phase timings on stderr
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    for (int i = 0; i < p; ++i)
        if (!isapprox(S(i, i), 1.0))                                 // group.jl:1932
            return set_error(ctx, KB_ERR_INVALID, "Σ must be scaled to a correlation matrix first.");
    if (!(threshold > 0.0 && threshold < 1.0))                       // group.jl:1933
        return set_error(ctx, KB_ERR_INVALID, "threshold should be in (0, 1) but was %g", threshold);
    // boundary case: linearly dependent columns (group.jl:1936-1958)
    std::vector<int> dependent;
    {
        std::vector<char> seen(p, 0);
        for (int i = 0; i < p; ++i)
            for (int j = i + 1; j < p; ++j) {
                // the predicate is "rows 1..i of columns i and j agree"; row i (Σ_ii = 1 against Σ_ij) is tested first
                // -- same boolean, but structured matrices whose columns share long equal prefixes (block-equicorrelated
                // Σ) no longer make this pre-pass O(p³)
                bool dep = isapprox(S(i, i), S(i, j));
                for (int k = 0; dep && k < i; ++k)
                    if (!isapprox(S(k, i), S(k, j))) dep = false;
                if (!dep) continue;
                int d = j;
                if (has_prior && std::find(prioritize.begin(), prioritize.end(), j) != prioritize.end()) d = i;
                if (!seen[d]) { seen[d] = 1; dependent.push_back(d); }   // unique!, first-appearance order
            }
    }
    if (!dependent.empty()) {                                        // group.jl:1960-1975
        std::vector<char> isdep(p, 0);
        for (int d : dependent) isdep[d] = 1;
        std::vector<int> indep;
        for (int i = 0; i < p; ++i)
            if (!isdep[i]) indep.push_back(i);
        const int q = (int)indep.size();
        std::vector<double> sub((size_t)q * q);
        std::vector<int64_t> gsub(q);
        for (int j = 0; j < q; ++j) {
            gsub[j] = groups[indep[j]];
            for (int i = 0; i < q; ++i) sub[i + (size_t)j * q] = S(indep[i], indep[j]);
        }
        std::vector<int> pr;
        if (has_prior)
            for (int idx : prioritize) {
                int c = 0;
                for (int d : dependent) c += (d < idx);
                pr.push_back(idx - c);
            }
        std::vector<int> rsub;
        KB_TRY(choose_group_reps_host(ctx, sub.data(), q, q, gsub.data(), threshold, has_prior, pr, &rsub));
        reps_out->clear();
        for (int i : rsub) reps_out->push_back(indep[i]);
        return KB_OK;
    }
    // all columns linearly independent: Σinv on the device (group.jl:1978), then g x g work per group
    const double t_prepass = since(t_start);
    const auto t_inv0 = std::chrono::steady_clock::now();
    std::vector<double> Sinv((size_t)p * p);
    {
        Scope sc(ctx);
        double *dS, *dI;
        KB_TRY(sc.alloc(&dS, (size_t)p * p));
        KB_TRY(sc.alloc(&dI, (size_t)p * p));
        KB_CUDA(ctx, cudaMemcpy2DAsync(dS, (size_t)p * sizeof(double), Sigma, (size_t)lds * sizeof(double),
                                       (size_t)p * sizeof(double), p, cudaMemcpyHostToDevice, ctx->stream));
        KB_TRY(spd_inverse_dev(ctx, dS, (long)p, p, dI, (long)p));
        KB_CUDA(ctx, cudaMemcpyAsync(Sinv.data(), dI, sizeof(double) * p * p, cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const double t_inv = since(t_inv0);
    const auto t_grp0 = std::chrono::steady_clock::now();
    SymView SI{Sinv.data(), (long)p};
    // unique(groups) in order of first appearance, members ascending
    std::unordered_map<int64_t, int> gid;
    std::vector<std::vector<int>> members;
    for (int i = 0; i < p; ++i) {
        auto it = gid.find(groups[i]);
        if (it == gid.end()) { it = gid.emplace(groups[i], (int)members.size()).first; members.emplace_back(); }
        members[it->second].push_back(i);
    }
    std::vector<int> reps;
    for (const auto& gidx : members) {
        const int gs = (int)gidx.size();
        if (gs == 1) { reps.push_back(gidx[0]); continue; }
        std::vector<int> index = select_best_rss_subset(gather_block(S, gidx, gidx), gs);   // group.jl:1993
        if (has_prior) {                                              // group.jl:1994-1998
            std::vector<int> first;
            for (int v : prioritize)
                for (size_t t = 0; t < index.size(); ++t)
                    if (gidx[index[t]] == v) { first.push_back((int)t); break; }
            index = prioritize_variants(index, first);
        }
        std::vector<int> idxS;
        for (int i : index) idxS.push_back(gidx[i]);
        std::vector<int> R{idxS[0]};
        reps.push_back(idxS[0]);
        for (int i = 1; i < gs && i < (int)idxS.size(); ++i) {
            std::vector<int> Rc, ROc;
            for (int v : idxS)
                if (std::find(R.begin(), R.end(), v) == R.end()) Rc.push_back(v);
            for (int v : gidx)
                if (std::find(R.begin(), R.end(), v) == R.end()) ROc.push_back(v);
            std::vector<double> RRinv = gather_block(S, R, R);
            if (!spd_inverse_small(RRinv, (int)R.size()))
                return set_error(ctx, KB_ERR_NOT_PD, "choose_group_reps: Σ[R, R] is not positive definite");
            std::vector<double> Q = gather_block(SI, ROc, ROc);       // inv(·) = Cov(ROc | RO)
            if (!spd_inverse_small(Q, (int)ROc.size()))
                return set_error(ctx, KB_ERR_NOT_PD, "choose_group_reps: inv(Σ)[ROc, ROc] is not positive definite");
            double ratio = 0.0;
            const int nr = (int)R.size(), nc = (int)ROc.size();
            for (int j : Rc) {
                double r2r = 0.0;                                     // Σ_Rj' inv(Σ_RR) Σ_Rj
                for (int a = 0; a < nr; ++a) {
                    double t = 0.0;
                    for (int b = 0; b < nr; ++b) t += RRinv[a + (size_t)b * nr] * S(R[b], j);
                    r2r += S(R[a], j) * t;
                }
                const int jj = (int)(std::find(ROc.begin(), ROc.end(), j) - ROc.begin());
                const double r2ro = S(j, j) - Q[jj + (size_t)jj * nc];   // Σ_ROj' inv(Σ_RORO) Σ_ROj
                ratio += r2r / r2ro;
            }
            ratio /= (double)Rc.size();
            if (ratio > threshold) break;                             // group.jl:2023-2024
            R.push_back(idxS[i]);
            reps.push_back(idxS[i]);
        }
    }
    std::sort(reps.begin(), reps.end());
    *reps_out = reps;
    if (trace)
        fprintf(stderr, "[choose_group_reps p=%d] dependent-column pre-pass %.3f s, inv(Sigma) on the device + D2H %.3f s, per-group "
                        "selection %.3f s\n", p, t_prepass, t_inv, since(t_grp0));
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------------
// cond_indep_corr (group.jl:205-240): Σnew on the device (full symmetric, ld p) from host Σ (upper trusted).
// ------------------------------------------------------------------------------------------------------
int cond_indep_corr_dev(kb_ctx* ctx, const double* Sigma, long lds, int p, const int64_t* groups, const std::vector<int>& reps,
                        double* dOut) {
    Perm P;
    KB_TRY(make_perm(ctx, p, reps, &P));
    const int r = (int)reps.size(), q = p - r;
    SymView S{Sigma, lds};
    // host: the group-block-diagonal pieces, in (reps, non-reps) coordinates
    std::vector<double> d12((size_t)r * std::max(q, 1), 0.0), b2((size_t)std::max(q, 1) * std::max(q, 1), 0.0);
    std::unordered_map<int64_t, std::pair<std::vector<int>, std::vector<int>>> by_group;   // positions in reps / in non
    for (int i = 0; i < r; ++i) by_group[groups[reps[i]]].first.push_back(i);
    for (int i = 0; i < q; ++i) by_group[groups[P.non[i]]].second.push_back(i);
    for (auto& kv : by_group) {
        const std::vector<int>&rp = kv.second.first, &np_ = kv.second.second;
        if (rp.empty())
            return set_error(ctx, KB_ERR_INVALID, "cond_indep_corr: group %lld has no representative", (long long)kv.first);
        if (np_.empty()) continue;
        std::vector<int> gr, gn;
        for (int i : rp) gr.push_back(reps[i]);
        for (int i : np_) gn.push_back(P.non[i]);
        const int a = (int)gr.size(), b = (int)gn.size();
        std::vector<double> inv_rr = gather_block(S, gr, gr);
        if (!spd_inverse_small(inv_rr, a)) return set_error(ctx, KB_ERR_NOT_PD, "cond_indep_corr: Σ[g_rep, g_rep] is not positive definite");
        std::vector<double> s_rc = gather_block(S, gr, gn);           // a x b
        std::vector<double> w((size_t)a * b);                         // inv(Σ_RR) Σ_RRc
        for (int j = 0; j < b; ++j)
            for (int i = 0; i < a; ++i) {
                double t = 0.0;
                for (int k = 0; k < a; ++k) t += inv_rr[i + (size_t)k * a] * s_rc[k + (size_t)j * a];
                w[i + (size_t)j * a] = t;
                d12[rp[i] + (size_t)np_[j] * r] = t;
            }
        for (int j = 0; j < b; ++j)
            for (int i = 0; i < b; ++i) {
                double t = 0.0;
                for (int k = 0; k < a; ++k) t += s_rc[k + (size_t)i * a] * w[k + (size_t)j * a];
                b2[np_[i] + (size_t)np_[j] * q] = S(gn[i], gn[j]) - t;
            }
    }
    Scope sc(ctx);
    double *dSig, *dP, *dd12, *db2, *dN12;
    int *dperm, *diperm;
    KB_TRY(sc.alloc(&dSig, (size_t)p * p));
    KB_TRY(sc.alloc(&dP, (size_t)p * p));
    KB_TRY(sc.alloc(&dd12, d12.size()));
    KB_TRY(sc.alloc(&db2, b2.size()));
    KB_TRY(sc.alloc(&dN12, (size_t)r * std::max(q, 1)));
    KB_TRY(sc.alloc(&dperm, (size_t)p));
    KB_TRY(sc.alloc(&diperm, (size_t)p));
    KB_CUDA(ctx, cudaMemcpy2DAsync(dSig, (size_t)p * sizeof(double), Sigma, (size_t)lds * sizeof(double),
                                   (size_t)p * sizeof(double), p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dperm, P.perm.data(), sizeof(int) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(diperm, P.iperm.data(), sizeof(int) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dd12, d12.data(), sizeof(double) * d12.size(), cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(db2, b2.data(), sizeof(double) * b2.size(), cudaMemcpyHostToDevice, ctx->stream));
    gather_sym_kernel<<<ewg(ctx, (long)p * p), 256, 0, ctx->stream>>>(dSig, (long)p, p, dperm, 1, dP, (long)p);
    ctx->launches++;
    if (q > 0) {
        // Σnew_12 = Σ11 Σ12_diag ;  Σnew_22 = Σblock2 + Σ12_diag' Σ11 Σ12_diag            (group.jl:231-238)
        KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, r, q, r, 1.0, dP, (long)p, dd12, (long)r, 0.0, dN12, (long)r));
        KB_TRY(gemm1(ctx, A_KCONTIG, B_KCONTIG, q, q, r, 1.0, dd12, (long)r, dN12, (long)r, 1.0, db2, (long)q));
    }
    scatter_blocks_kernel<<<ewg(ctx, (long)p * p), 256, 0, ctx->stream>>>(dP, (long)p, dN12, (long)r, db2, (long)std::max(q, 1), p, r,
                                                                          diperm, 0.0, dOut, (long)p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------------
// solve_s_graphical_group (group.jl:268-303).  dSig: device p x p FULL symmetric (ld p) -- Σ or cond_indep_corr(Σ);
// Sigma11: host r x r = Σ[reps, reps].  Outputs: S11_out (host r x r), dD (device p x p, full symmetric, |x| < 1e-10
// zeroed), obj.
// ------------------------------------------------------------------------------------------------------
int solve_s_graphical_group_dev(kb_ctx* ctx, const double* dSig, int p, const double* Sigma11, const int64_t* groups,
                                const std::vector<int>& reps, int method, double m, const kb_group_opts* opts, const double* V,
                                int nv, double* S11_out, double* dD, double* obj_out, kb_group_info* info) {
    Perm P;
    KB_TRY(make_perm(ctx, p, reps, &P));
    const int r = (int)reps.size(), q = p - r;
    std::vector<int64_t> g11(r);
    for (int i = 0; i < r; ++i) g11[i] = groups[reps[i]];
    // S, _, obj = solve_s_group(Symmetric(Σ11), groups[group_reps], method; m, kwargs...)     (group.jl:285-286)
    KB_TRY(solve_s_group_host(ctx, Sigma11, r, g11.data(), method, m, opts, V, nv, nullptr, nullptr, S11_out, obj_out, info));
    Scope sc(ctx);
    double *dP, *dS, *dInv, *dW, *dSW, *dT, *dD22;
    int *dperm, *diperm;
    const size_t rq = (size_t)r * std::max(q, 1), qq = (size_t)std::max(q, 1) * std::max(q, 1);
    KB_TRY(sc.alloc(&dP, (size_t)p * p));
    KB_TRY(sc.alloc(&dS, (size_t)r * r));
    KB_TRY(sc.alloc(&dInv, (size_t)r * r));
    KB_TRY(sc.alloc(&dW, rq));
    KB_TRY(sc.alloc(&dSW, rq));
    KB_TRY(sc.alloc(&dT, rq));
    KB_TRY(sc.alloc(&dD22, qq));
    KB_TRY(sc.alloc(&dperm, (size_t)p));
    KB_TRY(sc.alloc(&diperm, (size_t)p));
    KB_CUDA(ctx, cudaMemcpyAsync(dperm, P.perm.data(), sizeof(int) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(diperm, P.iperm.data(), sizeof(int) * p, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(ctx, cudaMemcpyAsync(dS, S11_out, sizeof(double) * r * r, cudaMemcpyHostToDevice, ctx->stream));
    gather_sym_kernel<<<ewg(ctx, (long)p * p), 256, 0, ctx->stream>>>(dSig, (long)p, p, dperm, 0, dP, (long)p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    if (q > 0) {
        const double* S12 = dP + (size_t)r * p;                      // Σ12 = P[0:r, r:p]
        const double* S22 = dP + (size_t)r * p + r;
        KB_TRY(spd_inverse_dev(ctx, dP, (long)p, r, dInv, (long)r));                                   // Σ11inv        (:289)
        KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, r, q, r, 1.0, dInv, (long)r, S12, (long)p, 0.0, dW, (long)r));    // Σ11inv Σ12 (:290)
        KB_TRY(gemm1(ctx, A_MCONTIG, B_KCONTIG, r, q, r, 1.0, dS, (long)r, dW, (long)r, 0.0, dSW, (long)r));      // S Σ11inv Σ12 (:291)
        sub_kernel<<<ewg(ctx, (long)r * q), 256, 0, ctx->stream>>>(dSW, (long)r, S12, (long)p, r, q, dT, (long)r);
        ctx->launches++;
        // D22 = Σ22 − Σ12' Σ11inv Σ12 + (Σ11inv Σ12)' S (Σ11inv Σ12) = Σ22 + W' (S W − Σ12)          (:296-297)
        GemmParams G{};
        G.M = q; G.N = q; G.nseg = 1;
        G.seg[0] = GemmSeg{dW, (long)r, dT, (long)r, r, 0};
        G.C = dD22; G.ldc = q; G.Cin = S22; G.ldcin = p; G.alpha = 1.0; G.beta = 1.0;
        KB_TRY(gemm(ctx, A_KCONTIG, B_KCONTIG, G));
    }
    // D in the original order, |x| < 1e-10 -> 0 (:300), symmetric from its upper triangle (Symmetric(D), :197)
    scatter_blocks_kernel<<<ewg(ctx, (long)p * p), 256, 0, ctx->stream>>>(dS, (long)r, dSW, (long)r, dD22, (long)std::max(q, 1), p, r,
                                                                          diperm, 1e-10, dD, (long)p);
    ctx->launches++;
    KB_CUDA(ctx, cudaGetLastError());
    KB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KB_OK;
}

}  // namespace kb
