"""Multi-GPU driver (run under torchrun on N GPUs): approx_modelX_gaussian_knockoffs with the windows dealt across
ranks (dist.approx_knockoffs_sharded; NCCL all-reduce of min(gamma) + the union of the s / mu segments).  Rank 0 also
runs the one-shot call on one GPU and checks that the union of the ranks' columns equals it bit for bit."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, p, w, m = int(os.environ.get("N", 4000)), int(os.environ.get("P", 4000)), int(os.environ.get("W", 500)), 2
S = kb.harness.ar1_sigma(p, 0.5)
X = np.asfortranarray(np.random.default_rng(1).standard_normal((n, p)) @ np.linalg.cholesky(S).T)
ctx = kb.Context(local)
dev = torch.device("cuda", local)
kb.dist.approx_knockoffs_sharded(ctx, X[:500, :w * world], "maxent", windowsize=w, m=m, seed=9, device=dev)   # warm-up
dist.barrier(); torch.cuda.synchronize()
t = time.time()
r = kb.dist.approx_knockoffs_sharded(ctx, X, "maxent", windowsize=w, m=m, seed=9, device=dev)
dist.barrier(); torch.cuda.synchronize()
wall = time.time() - t
c0, c1 = r["cols"]
mine = np.concatenate([r["Xko"][:, k * p + c0:k * p + c1] for k in range(m)], axis=1)
gathered = [None] * world
dist.all_gather_object(gathered, (rank, c0, c1, mine if os.environ.get("CHECK", "1") == "1" else None))
if rank == 0:
    print(f"approx sharded x{world}: n = {n}, p = {p}, windows of {w}, m = {m}: wall = {wall:.3f} s, gamma = {r['gamma']}, "
          f"column ranges = {[(g[1], g[2]) for g in gathered]}", flush=True)
    if os.environ.get("CHECK", "1") == "1":
        t = time.time()
        one = kb.approx_modelX_gaussian_knockoffs(X, "maxent", windowsize=w, m=m, seed=9, dense_sigma=False, ctx=ctx)
        t1 = time.time() - t
        ok = np.array_equal(one.s, r["s"])
        for _, a, b, cols in gathered:
            for k in range(m):
                ok = ok and np.array_equal(cols[:, k * (b - a):(k + 1) * (b - a)], one.Xko[:, k * p + a:k * p + b])
        print(f"single GPU one-shot: wall = {t1:.3f} s; union of the shards == one-shot bit for bit: {ok}", flush=True)
        assert ok
dist.barrier()
dist.destroy_process_group()
