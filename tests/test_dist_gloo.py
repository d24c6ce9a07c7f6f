"""CPU-only, world_size = 2 over gloo: the host-side logic of the multi-GPU path --
row sharding, the Gram stage's single all-reduce + local finish, round-robin placement of
independent solves.  The arithmetic inside each rank is numpy here; on GPUs the same drivers call the
`_dev` C-ABI entries (knockoffs.jl_b200/dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import knockoffs_b200  # noqa: F401
    kb = sys.modules["knockoffs_b200"]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, p = 1001, 17
        X = np.random.default_rng(3).standard_normal((n, p)) + 2.0
        lo, hi = kb.dist.shard_rows(n, world, rank)
        Xr = X[lo:hi]
        G = torch.from_numpy(Xr.T @ Xr)
        cs = torch.from_numpy(Xr.sum(0))
        kb.dist.allreduce_gram(G, cs)                    # the stage's only exchange
        Gn, csn = G.numpy(), cs.numpy()
        Sigma = (Gn - np.outer(csn, csn) / n) / n        # what kb_gram_finish_dev does on every rank
        Xc = X - X.mean(0)
        ok_gram = np.max(np.abs(Sigma - Xc.T @ Xc / n)) < 1e-10 and np.max(np.abs(csn / n - X.mean(0))) < 1e-12
        # independent solves dealt round-robin, results gathered on every rank
        res = kb.dist.solve_many(list(range(5)), lambda k: (k * k, rank))
        ok_many = [r[0] for r in res] == [0, 1, 4, 9, 16] and [r[1] for r in res] == [0, 1, 0, 1, 0]
        # row shards tile [0, n) exactly and the global-row keying makes shards independent of world size
        shards = [kb.dist.shard_rows(n, world, r) for r in range(world)]
        ok_shard = shards[0][0] == 0 and shards[-1][1] == n and all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
        # approx.jl windows dealt across ranks: the only exchange is min(γ) + the union of the s / μ̂ segments
        off = np.array([0, 40, 90, 100, 170, 200])
        parts = kb.dist.assign_windows(off, world)
        w0, w1 = parts[rank]
        c0, c1 = int(off[w0]), int(off[w1])
        s_full, mu_full = np.arange(200) + 1.0, -np.arange(200) - 0.5
        s, mu = np.zeros(200), np.zeros(200)
        s[c0:c1], mu[c0:c1] = s_full[c0:c1], mu_full[c0:c1]
        gam = kb.dist.merge_window_results(s, mu, [0.75, 0.5][rank], (c0, c1))
        ok_win = gam == 0.5 and np.array_equal(s, s_full) and np.array_equal(mu, mu_full) and parts[0][1] == parts[1][0]
        flag = torch.tensor([int(ok_gram and ok_many and ok_shard and ok_win)])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def test_sharded_gram_and_round_robin_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert q.get(timeout=5) == 1


def test_shard_rows_properties():
    sys.path.insert(0, ROOT)
    import knockoffs_b200  # noqa: F401
    kb = sys.modules["knockoffs_b200"]
    for n in (0, 1, 7, 100000, 500001):
        for w in (1, 2, 4, 8):
            sh = [kb.dist.shard_rows(n, w, r) for r in range(w)]
            assert sh[0][0] == 0 and sh[-1][1] == n
            sizes = [b - a for a, b in sh]
            assert max(sizes) - min(sizes) <= 1 and all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
    assert kb.dist.assign_round_robin(5, 2, 1) == [1, 3]
    G, mu = kb.dist.gram_reference_numpy([np.ones((2, 2)), 2 * np.ones((3, 2))])
    assert G.shape == (2, 2) and np.allclose(mu, 1.6)


def test_stagger_layout_and_fraction():
    """bench.py's multi-rank step: every row of the short method is sampled exactly once whatever the early fraction, the early
    ranks exclude the long solver, and the decision model reproduces the measured B200 regimes."""
    sys.path.insert(0, ROOT)
    import knockoffs_b200  # noqa: F401
    kb = sys.modules["knockoffs_b200"]
    for world in (1, 2, 3, 4, 8):
        for n in (1000, 1003, 100000):
            for f in (0.0, 0.25, 27.0 / 64.0, 1.0):
                cover = np.zeros(n, dtype=np.int64)
                for r in range(world):
                    e, l = kb.dist.stagger_layout(n, world, r, 1, f)
                    if r == 1 or world == 1:
                        assert e == (0, 0)                           # the long solver takes no early rows
                    for a, b in (e, l):
                        assert 0 <= a <= b <= n
                        cover[a:b] += 1
                assert np.all(cover == 1)
                if world > 1 and f == 1.0:
                    assert all(kb.dist.stagger_layout(n, world, r, 1, f)[1][0] == n for r in range(world))   # no late rows
    # measured stage times at p = 5000, n = 1e5, m = 5 (profiles/r02_bench_n*.json): A = 147, B = 510, Ts = 836 ms
    f2, plain2, stag2 = kb.dist.stagger_fraction(147.0, 510.0, 836.0, 2)
    assert 0.35 < f2 < 0.45 and stag2 < 0.9 * plain2
    assert kb.dist.stagger_fraction(115.0, 480.0, 836.0, 4)[0] == 1.0
    assert kb.dist.stagger_fraction(115.0, 480.0, 836.0, 8)[0] == 1.0
    fw, _, _ = kb.dist.stagger_fraction(115.0, 480.0, 8 * 836.0, 8)      # weak run: 8e5 rows in total
    assert 0.3 < fw < 0.4 and fw * 64 == int(fw * 64)
    assert kb.dist.stagger_fraction(500.0, 510.0, 836.0, 4)[0] == 0.0     # no gap: plain order
    assert kb.dist.stagger_fraction(147.0, 510.0, 836.0, 1) == (0.0, 510.0 + 2 * 836.0, 510.0 + 2 * 836.0)
