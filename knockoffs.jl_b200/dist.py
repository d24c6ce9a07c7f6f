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


def gram_reference_numpy(X_shards: Sequence[np.ndarray], center=True):
    """Host restatement of the sharded Gram algebra (used by the gloo tests to check the partition logic)."""
    n = sum(x.shape[0] for x in X_shards)
    G = sum(x.T @ x for x in X_shards)
    cs = sum(x.sum(0) for x in X_shards)
    if center:
        G = G - np.outer(cs, cs) / n
    return G / n, cs / n
