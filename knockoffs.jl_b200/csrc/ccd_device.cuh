// Device helpers shared by the CCD kernels (ccd_stream.cu, ccd_block.cu).
#pragma once
#include "ccd.cuh"

namespace kb {

constexpr int SLOT_ROWS = CCD_SLOT_ROWS;
constexpr double SKIP_ABS = 1e-40;    // slot updates with |c m_k| max|m_rb| below this are dropped (banded Σ)
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000LL;   // ~2 s: watchdog, never reached in a correct run

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct SlotPos {
    int k;    // column
    int rb;   // 64-row block
};

// slot enumeration: column k owns row blocks rb = k/64 .. nrb-1
__device__ __forceinline__ SlotPos slot_from_index(long idx, int nrb) {
    // columns come in groups of 64 with the same slot count (nrb - g); walk groups (<= p/64 steps, once)
    int g = 0;
    long per = (long)SLOT_ROWS * nrb;
    while (g < nrb - 1 && idx >= per) {
        idx -= per;
        ++g;
        per = (long)SLOT_ROWS * (nrb - g);
    }
    const int cnt = nrb - g;
    SlotPos s;
    s.k = g * SLOT_ROWS + (int)(idx / cnt);
    s.rb = g + (int)(idx % cnt);
    return s;
}

__host__ __device__ __forceinline__ long slot_index(int k, int rb, int nrb) {
    const int g = k / SLOT_ROWS;
    // slots before group g: 64 * sum_{h<g} (nrb - h)
    const long before = (long)SLOT_ROWS * ((long)g * nrb - (long)g * (g - 1) / 2);
    return before + (long)(k - g * SLOT_ROWS) * (nrb - g) + (rb - g);
}

// deterministic block reduction of (sum) -- identical instruction sequence in every CTA, so every
// CTA derives bit-identical line-search scalars from the same published column.
template <int CCD_WARPS>
__device__ __forceinline__ double block_sum(double v, double* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < CCD_WARPS; ++w) t += red[w];
    __syncthreads();
    return t;
}


// Reciprocal / division / square root for the SERIAL chain of the blocked kernel (ccd_block_seq_kernel walks 64 line searches
// one after the other; ncu r02: 'wait' = fixed-latency dependency is its top stall).  The IEEE sequences the compiler emits
// are ~15-25 dependent instructions each plus a slow-path test; these are a MUFU seed + two Newton steps + one residual
// correction (<= 1 ulp, not correctly rounded), about a third of the chain.  Inf / NaN / zero inputs still produce
// Inf / NaN, so the reference's error paths (utilities.jl:195-196) are unchanged.
__device__ __forceinline__ double chain_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
template <bool FAST>
__device__ __forceinline__ double chain_div(double a, double b) {
    if (!FAST) return a / b;
    const double r = chain_rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
template <bool FAST>
__device__ __forceinline__ double chain_sqrt(double x) {
    if (!FAST) return sqrt(x);
    if (!(x > 0.0)) return sqrt(x);                      // 0, negative, NaN: the library's result
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    const double s = x * r;
    return fma(fma(-s, s, x), 0.5 * r, s);
}

// One coordinate's line search -- identical arithmetic in every thread of every CTA.
//   MVR      utilities.jl:149-159 + solve_quadratic :187-199   (cn = -ss, cd = Mjj)
//   maxent   utilities.jl:243-264  with ‖L⁻¹ỹ‖² = A_jj (A_jj M_jj − 1); s takes the UNCLIPPED value (quirk Q2)
//   sdp_ccd  utilities.jl:318-336  (δ = s_j − s_new, sign flipped)
// Outputs: delta (A <- A − delta e_j e_j'), s_new, skip (|δ| < 1e-15), cfac = δ / (1 − δ M_jj) (0 when skipped),
// status (KB_ERR_NAN / KB_ERR_NOT_PD on the reference's error paths).
template <bool FAST = false>
__device__ __forceinline__ void ccd_line_search(int method, double m, double lam, double lambda_min, double sj, double ajj,
                                                double kSjj, double ss, double Mjj, double& delta, double& s_new, bool& skip,
                                                double& cfac, int& status) {
    delta = 0.0;
    s_new = sj;
    skip = false;
    if (method == KB_METHOD_MVR) {
        const double cn = -ss, cd = Mjj;
        const double qa = -cn - cd * cd * m * m;
        const double qb = 2.0 * (-cn * sj + cd * m * m);
        const double qc = -cn * sj * sj - m * m;
        double dj;
        if (qa == 0.0 && qc == 0.0) {
            dj = 0.0;
        } else {
            const double disc = chain_sqrt<FAST>(qb * qb - 4.0 * qa * qc);
            const double inv2a = FAST ? chain_rcp(2.0 * qa) : 0.0;
            const double x1 = FAST ? (-qb + disc) * inv2a : (-qb + disc) / (2.0 * qa);
            const double x2 = FAST ? (-qb - disc) * inv2a : (-qb - disc) / (2.0 * qa);
            dj = (-sj < x1 && x1 < chain_div<FAST>(1.0, cd)) ? x1 : x2;
            if (isinf(dj) || isnan(dj)) status = KB_ERR_NAN;
        }
        const double ub = chain_div<FAST>(1.0, cd) - lambda_min;
        if (dj > ub) dj = ub;
        if (fabs(dj) < 1e-15) skip = true;
        else { s_new = sj + dj; delta = dj; }
    } else {
        const double x_l2 = ajj * (ajj * Mjj - 1.0);
        const double zeta = kSjj - sj;
        const double c = chain_div<FAST>(zeta * x_l2, zeta + x_l2);
        if (method == KB_METHOD_MAXENT) {
            const double sn = (kSjj - c) / 2.0;
            const double ub = chain_div<FAST>(1.0, Mjj) - lambda_min;
            double d = sn - sj;
            if (d > ub) d = ub;
            if (fabs(d) < 1e-15) skip = true;
            else { s_new = sn; delta = d; }
        } else {
            double sn = kSjj - c - lam;
            sn = sn < 0.0 ? 0.0 : (sn > 1.0 ? 1.0 : sn);
            const double d = sj - sn;
            if (fabs(d) < 1e-15) skip = true;
            else { s_new = sn; delta = -d; }
            if (isnan(sn)) status = KB_ERR_NAN;
        }
    }
    cfac = 0.0;
    if (!skip && status == 0) {
        const double den = 1.0 - delta * Mjj;
        if (!(den > 0.0)) status = KB_ERR_NOT_PD;        // the downdate would leave the PD cone (utilities.jl:591-593)
        else cfac = chain_div<FAST>(delta, den);
    }
    if (status != 0) cfac = 0.0;
}

}  // namespace kb
