// Representative group knockoffs (src/group.jl:170-303, 1929-2033) -- see rep.cu.
#pragma once
#include "kb_common.cuh"
#include <vector>

namespace kb {

int choose_group_reps_host(kb_ctx* ctx, const double* Sigma, long lds, int p, const int64_t* groups, double threshold,
                           bool has_prior, const std::vector<int>& prioritize, std::vector<int>* reps_out);

int cond_indep_corr_dev(kb_ctx* ctx, const double* Sigma, long lds, int p, const int64_t* groups, const std::vector<int>& reps,
                        double* dOut);

int solve_s_graphical_group_dev(kb_ctx* ctx, const double* dSig, int p, const double* Sigma11, const int64_t* groups,
                                const std::vector<int>& reps, int method, double m, const kb_group_opts* opts, const double* V,
                                int nv, double* S11_out, double* dD, double* obj_out, kb_group_info* info);

}  // namespace kb
