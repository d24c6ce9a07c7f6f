"""Device-pointer (`_dev`) entry points: the same stages on buffers that already live in HBM.
Pointers are plain integers (e.g. ``tensor.data_ptr()``); matrices are column-major float64.
torch is used by callers for allocation and torch.distributed only -- never for the arithmetic."""
from __future__ import annotations

import ctypes as C

from ._lib import SolveInfo, check
from .api import Context, _info_dict, _method_id, make_opts


def solve_s_dev(ctx: Context, dSigma: int, p: int, method, m, d_s_out: int, d_s_init: int = 0, **kwargs) -> dict:
    """kb_solve_s_dev: Sigma (p x p, ld p, upper triangle trusted) and s live on ctx's device."""
    opts = make_opts(**kwargs)
    info = SolveInfo()
    rc = ctx._lib.kb_solve_s_dev(ctx.h, dSigma, p, _method_id(method), float(m), C.byref(opts), d_s_init or None,
                                 d_s_out, C.byref(info))
    check(ctx.h, rc)
    return _info_dict(info, str(method).lstrip(":"))


def cond_factor_block_dev(ctx: Context, dSigma: int, p: int, ds: int, k: int, m: int, dSiS: int, dLkk: int, dMk: int):
    """Block column k (1-based) of the factor for Diagonal(s) from the closed form of the recursion (pointers; 0 = NULL)."""
    check(ctx.h, ctx._lib.kb_cond_factor_block_dev(ctx.h, dSigma, p, ds, int(k), int(m), dSiS or None, dLkk, dMk or None))


def cond_factor_dev(ctx: Context, dSigma: int, p: int, dS: int, s_is_diag: bool, m: int, dSiS: int, dLb: int):
    check(ctx.h, ctx._lib.kb_cond_factor_dev(ctx.h, dSigma, p, dS, int(s_is_diag), int(m), dSiS, dLb))


def sample_dev(ctx: Context, dX: int, n: int, ldx: int, p: int, dmu: int, dSiS: int, dLb: int, m: int, dXko: int,
               ldxko: int, dZ: int = 0, ldz: int = 0, seed: int = 0, row_offset: int = 0):
    check(ctx.h, ctx._lib.kb_sample_dev(ctx.h, dX, n, ldx, p, dmu, dSiS, dLb, int(m), dZ or None, ldz, int(seed),
                                        int(row_offset), dXko, ldxko))


def gram_partial_dev(ctx: Context, dX: int, n_local: int, ldx: int, p: int, dG: int, dcolsum: int):
    check(ctx.h, ctx._lib.kb_gram_partial_dev(ctx.h, dX, n_local, ldx, p, dG, dcolsum))


def gram_finish_dev(ctx: Context, dG: int, dcolsum: int, n_total: int, p: int, center: bool = True, denom: float = 0.0):
    check(ctx.h, ctx._lib.kb_gram_finish_dev(ctx.h, dG, dcolsum, n_total, p, int(center), float(denom)))
