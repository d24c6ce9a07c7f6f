"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in
the CPU tests).  Only the naturally data-parallel stages are partitioned:

* Gram: rows of X are sharded; each rank computes X_r'X_r and 1'X_r on its own GPU, ONE all-reduce of
  the p x p sums (+ the p column sums) over NVLink, then every rank finishes (centre, scale) locally.
* Sampling: rows are sharded; the factor is replicated; no collective at all -- the device RNG is
  keyed by (seed, GLOBAL row, column), so the union of the shards equals the single-GPU result.
* A single CCD solve is sequential in its coordinates and stays on one GPU ("replicas only"):
  independent solves (different methods / m / windows) are dealt round-robin to ranks and the
  results all-gathered.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def shard_rows(n: int, world: int, rank: int):
    """Contiguous row shard [lo, hi) of rank `rank`; remainders go to the first ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def assign_round_robin(n_items: int, world: int, rank: int) -> List[int]:
    return list(range(rank, n_items, world))


# ---- two solves of very different length on different ranks (bench.py: :mvr on rank 0, :sdp_ccd on rank 1) ------------
# While the long solve is still running, the other ranks ("early" ranks) already hold the short method's factor and sample a
# fraction f of ITS rows among themselves; afterwards all ranks share the remaining (1 - f) n rows of the short method and the
# n rows of the long one.  No data-path collective is added: the device RNG is keyed by the global row, so the union of the
# pieces is the single-GPU result whatever the split.
def stagger_fraction(A_ms: float, B_ms: float, Ts_ms: float, world: int, bcast_ms: float = 5.0, late_bcast_ms: float = 4.0):
    """f in [0, 1] and the modelled step times (plain, staggered).  A / B: solve + factor time of the short / long method on
    their ranks; Ts: time to sample ALL rows of one method on one GPU.  plain = B + 2 Ts / N;  the early ranks can sample
    f = (B - A - bcast)(N - 1) / Ts of the short method's rows before the long solve ends;  staggered = max(A + f Ts / (N - 1), B)
    + (2 - f) Ts / N (+ one more broadcast of the short method's blocks to the long solver when f < 1).  f is cut to 1/64ths
    so that every rank derives the same row counts; 0 when staggering is not worth at least 2 %."""
    plain = B_ms + 2.0 * Ts_ms / world
    if world < 2 or Ts_ms <= 0.0:
        return 0.0, plain, plain
    gap = B_ms - A_ms - bcast_ms
    f = min(1.0, 0.97 * gap * (world - 1) / Ts_ms) if gap > 0.0 else 0.0
    f = float(int(f * 64.0)) / 64.0 if f < 1.0 else 1.0
    stag = max(A_ms + f * Ts_ms / (world - 1), B_ms) + (2.0 - f) * Ts_ms / world + (late_bcast_ms if f < 1.0 else 0.0)
    if f < 0.1 or stag >= 0.98 * plain:
        return 0.0, plain, stag
    return f, plain, stag


def stagger_layout(n_total: int, world: int, rank: int, long_rank: int, f: float):
    """Row ranges of the SHORT method for `rank`: (early range, late range).  The first round(f n) rows are dealt to the ranks
    other than `long_rank` (early), the rest to all ranks (late).  The long method's rows are the plain shard_rows()."""
    n_early = int(round(f * n_total)) if world > 1 else 0
    early_ranks = [r for r in range(world) if r != long_rank]
    if n_early > 0 and rank in early_ranks:
        e_rng = shard_rows(n_early, len(early_ranks), early_ranks.index(rank))
    else:
        e_rng = (0, 0)
    l0, l1 = shard_rows(n_total - n_early, world, rank)
    return e_rng, (l0 + n_early, l1 + n_early)


def allreduce_gram(G, colsum, group=None):
    """The Gram stage's only exchange: sum of the per-rank p x p partial Grams and column sums."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
    return G, colsum


def gram_sharded_dev(ctx, X_local, n_total: int, center: bool = True, denom: float = 0.0, group=None):
    """X_local: torch float64 CUDA tensor holding this rank's rows COLUMN-major, i.e. a (p, n_local)
    contiguous tensor whose [j, i] is X[i, j].  Returns (Sigma (p x p), colmean (p)) replicated on every rank."""
    import torch
    from . import dev
    p, n_local = X_local.shape
    G = torch.empty((p, p), dtype=torch.float64, device=X_local.device)
    cs = torch.empty(p, dtype=torch.float64, device=X_local.device)
    torch.cuda.current_stream().synchronize()
    dev.gram_partial_dev(ctx, X_local.data_ptr(), n_local, n_local, p, G.data_ptr(), cs.data_ptr())
    allreduce_gram(G, cs, group)
    torch.cuda.current_stream().synchronize()
    dev.gram_finish_dev(ctx, G.data_ptr(), cs.data_ptr(), n_total, p, center, denom)
    return G, cs


def solve_many(problems: Sequence, solve_one: Callable, group=None):
    """Independent solves dealt round-robin to ranks, results gathered on every rank
    (SURVEY 8e 'replicas only': e.g. :mvr and :sdp_ccd of config 3, or approx.jl windows)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = {i: solve_one(problems[i]) for i in assign_round_robin(len(problems), world, rank)}
    if world == 1:
        return [mine[i] for i in range(len(problems))]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    out = {}
    for d in gathered:
        out.update(d)
    return [out[i] for i in range(len(problems))]


def torch_reduce_hook(group=None, device=None):
    """reduce(op, values) for api.solve_s_group_sharded backed by torch.distributed (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    ops = {0: dist.ReduceOp.MIN, 1: dist.ReduceOp.SUM, 2: dist.ReduceOp.MAX}

    def reduce(op, values):
        t = torch.from_numpy(values.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=ops[op], group=group)
        values[:] = t.cpu().numpy()
    return reduce


def solve_s_group_blockdiag(ctx, Sigma_blocks, groups_blocks, method, m=1, group=None, device=None, **kwargs):
    """Block-diagonal Σ = blockdiag(Sigma_blocks[0], Sigma_blocks[1], ...) with one (or more) diagonal block per
    rank: this rank solves `Sigma_blocks[rank]` in lockstep with the others.  Returns
    (S_block_of_this_rank, total_objective)."""
    import torch
    import torch.distributed as dist
    from . import api
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    hook = torch_reduce_hook(group, device) if dist.is_initialized() and dist.get_world_size(group) > 1 \
        else (lambda op, values: None)
    S, obj, info = api.solve_s_group_sharded(Sigma_blocks[rank], groups_blocks[rank], method, hook, m=m, ctx=ctx, **kwargs)
    tot = np.array([obj])
    hook(1, tot)
    return S, float(tot[0]), info


def assign_windows(offsets: Sequence[int], world: int) -> List[tuple]:
    """Contiguous window ranges [w0, w1) per rank for approx.jl's independent windows, balanced on the windows'
    cubic solve cost (Σ pw³): rank r gets the windows whose cumulative-cost midpoint falls in its 1/world share."""
    pw = np.diff(np.asarray(offsets, dtype=np.int64)).astype(np.float64)
    cost = pw ** 3
    cum = np.cumsum(cost)
    mid = (cum - 0.5 * cost) / max(cum[-1], 1e-300)
    owner = np.minimum((mid * world).astype(int), world - 1)
    out = []
    for r in range(world):
        idx = np.flatnonzero(owner == r)
        out.append((int(idx[0]), int(idx[-1]) + 1) if idx.size else (0, 0))
    return out


def merge_window_results(s, mu, gamma_local, cols, group=None, device=None):
    """The approx path's only exchange: min of the ranks' γ (approx.jl:97 on a block-diagonal Σ) and the union of
    their s / μ̂ segments (each rank owns the columns cols = [c0, c1)).  In place on the numpy arrays; returns γ."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(gamma_local)
    c0, c1 = cols
    buf = np.zeros((2, s.shape[0]))
    buf[0, c0:c1], buf[1, c0:c1] = s[c0:c1], mu[c0:c1]
    t = torch.from_numpy(buf)
    g = torch.tensor([float(gamma_local)], dtype=torch.float64)
    if device is not None:
        t, g = t.to(device), g.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)         # disjoint supports: sum = union
    dist.all_reduce(g, op=dist.ReduceOp.MIN, group=group)
    t = t.cpu().numpy()
    s[:], mu[:] = t[0], t[1]
    return float(g.item())


def approx_knockoffs_sharded(ctx, X, method, window_ranges=None, m=1, windowsize=500, Z=None, seed=0, group=None,
                             device=None, **kwargs):
    """approx_modelX_gaussian_knockoffs (src/approx.jl:62-102) with the windows dealt across ranks: every rank
    estimates Σ̂_w and solves s_w for its contiguous window range, ONE exchange (min γ, union of s and μ̂), then
    every rank samples the X̃ columns of its own windows.  Returns dict(cols=(c0, c1), Xko (n x m p, only this
    rank's columns of every copy are filled), s, mu, gamma, blocks_packed, offsets)."""
    import ctypes as C
    import torch.distributed as dist
    from . import api
    from ._lib import SolveInfo, check
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    X = api._f(X)
    n, p = X.shape
    off, _ = api._window_offsets(p, window_ranges, windowsize)
    nwin = len(off) - 1
    w0, w1 = assign_windows(off, world)[rank]
    c0, c1 = int(off[w0]), int(off[w1])
    opts = api.make_opts(**kwargs)
    info = SolveInfo()
    packed = np.zeros(int(np.sum(np.diff(off) ** 2)))
    s, mu = np.zeros(p), np.zeros(p)
    gam = C.c_double(1.0)
    check(ctx.h, ctx._lib.kb_approx_solve_windows(ctx.h, api._ptr(X), n, p, api._ptr(off), nwin, w0, w1,
                                                  api._method_id(method), int(m), C.byref(opts), api._ptr(packed),
                                                  api._ptr(s), api._ptr(mu), C.byref(gam), C.byref(info)))
    gamma = merge_window_results(s, mu, gam.value, (c0, c1), group, device)
    s *= gamma                                                     # approx.jl:98
    Xko = np.empty((n, int(m) * p), order="F")
    Zf = None if Z is None else api._f(Z)
    check(ctx.h, ctx._lib.kb_approx_condition_windows(ctx.h, api._ptr(X), n, p, api._ptr(off), nwin, w0, w1, int(m),
                                                      api._ptr(packed), api._ptr(s), api._ptr(mu), api._ptr(Zf), int(seed),
                                                      0, api._ptr(Xko)))
    return dict(cols=(c0, c1), Xko=Xko, s=s, mu=mu, gamma=gamma, blocks_packed=packed, offsets=off)


def gram_reference_numpy(X_shards: Sequence[np.ndarray], center=True):
    """Host restatement of the sharded Gram algebra (used by the gloo tests to check the partition logic)."""
    n = sum(x.shape[0] for x in X_shards)
    G = sum(x.T @ x for x in X_shards)
    cs = sum(x.sum(0) for x in X_shards)
    if center:
        G = G - np.outer(cs, cs) / n
    return G / n, cs / n
