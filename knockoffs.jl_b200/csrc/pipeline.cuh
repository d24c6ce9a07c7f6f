// Device-side stages (2)-(4): Gram, conditional-covariance factor, fused sampling GEMM, RNG.
#pragma once
#include "kb_common.cuh"

namespace kb {

void default_solve_opts(kb_solve_opts* o);

int solve_s_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, int method, double m, const kb_solve_opts* opts,
                const double* d_s_init, double* d_s_out, kb_solve_info* info);

// inverse_iteration: place the probes with Rayleigh-quotient upper bounds (solve.cu).  Off for the group solver, whose
// sharded form needs every rank to walk the SAME deterministic probe sequence as the single-GPU solve.
int eigmin_lower_dev(kb_ctx* ctx, const double* B, long ld, int p, double* work_A, double* invdiag, int* d_fail,
                     double* lambda_out, bool inverse_iteration = false);
// λmin(Symmetric(Sigma)) for a matrix given by its upper triangle
int eigmin_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, double* lambda_out);

// dSigma: upper triangle trusted, leading dim lds.  dS: p-vector (diag) or p x p (ld = ldS).
// Outputs: dSiS (p x p, ld ldo) and dLb (2m-1 blocks of p x p, each ld ldo, block stride ldo*p).
int cond_factor_block_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* ds, int k, int m,
                          double* dSiS, double* dLkk, double* dMk, long ldo);
int cond_factor_dev(kb_ctx* ctx, const double* dSigma, long lds, int p, const double* dS, long ldS, int s_is_diag,
                    int m, double* dSiS, double* dLb, long ldo);

// TMA tensor maps of the structured sampling launch (dgemm_tma.cu); opaque here (4 CUtensorMap objects)
struct alignas(64) SampleTmaMaps {
    unsigned char raw[4 * 128];
};
bool sample_tma_usable(const double* T, long ldt, long tstride, const double* L, long ldl, long lstride, const double* Nn, long ldn,
                       long nstride);
int sample_tma_maps(kb_ctx* ctx, SampleTmaMaps* out, const double* T, long chunk, int p, int m, long tstride, const double* Lb, long ldo,
                    long bstride, const double* Nn, long ldn);
int sample_tma_launch(kb_ctx* ctx, const SampleTmaMaps& mp, int rows, int p, int m, int band, const double* Mu, long ldmu, double* Xko,
                      long ldxko);

// Row-chunk sampling plan: the per-call constants (−ΣinvS, μ'ΣinvS) and the chunk work buffers.
struct SamplePlan {
    int p = 0, m = 1;
    long chunk = 0;            // max rows per sample_chunk call (even)
    long ld = 0;               // leading dimension of nSiS
    double* nSiS = nullptr;    // −ΣinvS
    double* bias = nullptr;    // μ'ΣinvS
    double* Mu = nullptr;      // m > 1: X − (X .− μ')ΣinvS of the current chunk
    double* T = nullptr;       // m > 1: suffix sums of the Z blocks (m − 1 exclusive ones, or m inclusive ones if structured)
    bool structured = false;   // every M_k = L_kk + N_k with N_k upper triangular up to a narrow band below the diagonal:
                               // copies cost 2 n p² instead of 3 n p²
    int band = 0;              // rows below the diagonal where N_k may be non-zero (0 for Diagonal(s); g − 1 for group blocks)
    double* N = nullptr;       // structured: N_k = triu(M_k) − diag(L_kk), m − 1 blocks of ld x p
    double* Zbuf = nullptr;    // device-generated Z of the current chunk (chunk x m p)
    const double* dLb = nullptr;
    long ldo = 0;
    bool tma = false;          // structured launch through the TMA / mbarrier kernel (sample_tma_kernel)
    SampleTmaMaps maps;
};
int sample_plan_create(kb_ctx* ctx, Scope& sc, int p, int m, const double* dmu, const double* dSiS, const double* dLb,
                       long ldo, long chunk, bool need_zbuf, SamplePlan* plan);
// enqueue (no synchronisation) rows [0, rows) of one chunk; Zc == NULL draws Z from the Philox stream
// for global rows row_global0 ..; Xc / Zc / Xko are chunk-local pointers with their leading dimensions.
int sample_chunk(kb_ctx* ctx, const SamplePlan& plan, const double* Xc, long rows, long ldx, const double* Zc, long ldz,
                 uint64_t seed, int64_t row_global0, double* Xko, long ldxko);
// Fused sampling on device-resident row slabs (rows [0,n) of dX/dZ/dXko with the given leading dims).
int sample_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, const double* dmu, const double* dSiS,
               const double* dLb, long ldo, int m, const double* dZ, long ldz, uint64_t seed, int64_t row_offset,
               double* dXko, long ldxko);

// band of the factor blocks' structure as sample_plan_create detects it (device blocks, ld ldo): -1 = general form
int factor_band_dev(kb_ctx* ctx, const double* dLb, long ldo, int p, int m, int* band_out);

int randn_dev(kb_ctx* ctx, uint64_t seed, int64_t row_offset, long n, long ncols, long col0, double* dZ, long ldz);

int gram_partial_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, double* dG, long ldg, double* dcolsum,
                     int accumulate);
int gram_finish_dev(kb_ctx* ctx, double* dG, long ldg, double* dcolsum, long n_total, int p, int center, double denom);

// Ledoit-Wolf shrinkage pieces (pipeline.cu)
int gram_sq_partial_dev(kb_ctx* ctx, const double* dX, long n, long ldx, int p, const double* dmean, double* dY, long ldy,
                        double* dG2, long ldg, int accumulate);
int lw_finish_dev(kb_ctx* ctx, double* dS, long lds, const double* dG2, long ldg, int p, long n_total, double* lambda_out);

// dense helpers on matrices given by their upper triangle (Symmetric(.), uplo = :U)
int spd_inverse_dev(kb_ctx* ctx, const double* dA_upper, long lda, int p, double* dOut, long ldo);
int potrf_dev(kb_ctx* ctx, const double* dA_upper, long lda, int p, double* dL, long ldl);
int sym_upper_to_full_dev(kb_ctx* ctx, const double* dA, long lda, int p, double* dOut, long ldo);

}  // namespace kb
