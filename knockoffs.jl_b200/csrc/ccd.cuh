// Shared declarations of the CCD s-solver kernels.
#pragma once
#include "kb_common.cuh"

namespace kb {

constexpr int CCD_SLOT_ROWS = 64;   // rows per slot; the M buffer's leading dimension is a multiple of this
constexpr int CCD_BLOCK_WIDTH = 64; // coordinates per delayed-update block (ccd_block.cu)

struct CcdParams {
    int p;
    int method;           // KB_METHOD_MVR / MAXENT / SDP_CCD
    double m;
    int nsweeps;          // sweeps to run in this launch
    int sweep0;           // index of the first sweep of this launch (trace slot)
    double tol, lambda_min;
    double sdp_lambda;    // barrier coefficient at the first sweep of this launch
    double sdp_mu;
    double* M;            // A^{-1}: column-major, lower triangle + symmetric 64x64 diagonal blocks valid,
    long ld;              //   leading dimension ld (multiple of 64), rows p..ld-1 are zero
    double* s;            // current s (correlation scale)
    double* adiag;        // diagonal of the matrix A whose inverse M is (tracks clipped updates)
    const double* ksig_diag;   // kappa * Sigma_jj
    double* colbuf;       // 3 column buffers (see ccd_stream_colbuf_doubles)
    unsigned long long* ready;   // [0] publish counter, [1] abort flag
    int* status;
    int* sweeps_out;      // sweeps executed by this launch
    double* trace;        // KB_MAX_TRACE per-sweep max|delta|
    double* counters;     // [0] slots streamed, [1] coordinates updated, [2] sdp lambda after the launch, [3] max|delta| of its last sweep
};

size_t ccd_stream_colbuf_doubles(int p);
// leading dimension the streaming kernel needs for a p x p inverse
static inline long ccd_stream_ld(int p) { return ((long)p + CCD_SLOT_ROWS - 1) / CCD_SLOT_ROWS * CCD_SLOT_ROWS; }
// one launch = init kernel + persistent cooperative kernel running P.nsweeps sweeps
int ccd_stream_run(kb_ctx* ctx, const CcdParams& P, int* nctas_out);

// Blocked (delayed rank-64 update) variant, ccd_block.cu: ONE sweep per call on the same resident inverse.
size_t ccd_block_work_doubles(int p);
int ccd_block_run(kb_ctx* ctx, const CcdParams& P, double* work);

}  // namespace kb
