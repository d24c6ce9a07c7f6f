"""knockoffs_b200 -- B200-native (sm_100a) model-X Gaussian knockoff hot path behind the Knockoffs.jl API.

The directory is named ``knockoffs.jl_b200`` (not importable by name because of the dot); import it as
``knockoffs_b200`` through the loader module at the repo root.
"""
from ._lib import KnockoffsError, PosDefException, LIB_PATH, SIGNATURES, load  # noqa: F401
from .api import (Context, GaussianGroupKnockoff, GaussianKnockoff, make_group_opts, modelX_gaussian_group_knockoffs,
                  solve_s_group, solve_s_group_sharded, cond_factor, condition, cov_shrink_lw, default_context, dense_factor_from_blocks,  # noqa: F401
                  eigmin, gram, hook_dgemm, hook_factor_band, hook_potrf, hook_potrf_positive, hook_randn, hook_spd_inverse, make_opts,
                  modelX_gaussian_knockoffs, sample, solve_equi, solve_s, threshold, mk_threshold, MK_statistics,
                  marginal_stats, fit_marginal, MarginalKnockoffFilter, approx_modelX_gaussian_knockoffs,
                  ApproxGaussianKnockoff, feasible_gamma, choose_group_reps, cond_indep_corr,
                  solve_s_graphical_group, modelX_gaussian_rep_group_knockoffs, GaussianRepGroupKnockoff, MultiContext)
from . import harness, dev, dist  # noqa: F401

__all__ = ["Context", "GaussianKnockoff", "KnockoffsError", "PosDefException", "cond_factor", "condition", "eigmin",
           "gram", "modelX_gaussian_knockoffs", "sample", "solve_equi", "solve_s", "harness"]
