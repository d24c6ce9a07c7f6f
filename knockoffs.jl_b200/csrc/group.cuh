// Group knockoff solver entry points (group.cu).
#pragma once
#include "kb_common.cuh"
#include <vector>

namespace kb {

void default_group_opts(kb_group_opts* o);

// Σcor: host p x p column-major correlation matrix whose groups are contiguous (offsets goff, ng + 1 entries).
// Vperm (optional): p x nv directions in the same (contiguous) ordering, each supported on one group.
// Output: the group blocks of S (ng blocks, stride gmax², ld gmax) and the accumulated objective.
int group_solve_contiguous(kb_ctx* ctx, const std::vector<double>& Scor, int p, const std::vector<int>& goff, int method,
                           double m, const kb_group_opts& opts, const double* Vperm, int nv, kb_reduce_fn reduce,
                           void* user, std::vector<double>& Sblocks_out, int* gmax_out, double* obj_out,
                           kb_group_info* info);

// solve_s_group (group.jl:348-422) on host inputs: cov->cor, sort groups contiguous, dispatch, un-permute, cor->cov.
int solve_s_group_host(kb_ctx* ctx, const double* Sigma, int p, const int64_t* groups, int method, double m,
                       const kb_group_opts* opts, const double* V, int nv, kb_reduce_fn reduce, void* user,
                       double* S_out, double* obj_out, kb_group_info* info);

}  // namespace kb
