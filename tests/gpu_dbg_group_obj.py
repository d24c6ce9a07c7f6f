import os, sys, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb
from oracle import group_oracle as go
pb = 250
blocks, groups = [], []
for b in range(2):
    g = np.repeat(np.arange(1, pb // 5 + 1), 5)
    rho = 0.7 + 0.02 * b
    S = np.where(g[:, None] == g[None, :], rho, rho * 0.25).astype(float); np.fill_diagonal(S, 1.0)
    blocks.append(np.asfortranarray(S)); groups.append(g)
n = 2 * pb
full = np.zeros((n, n), order="F"); gfull = np.zeros(n, dtype=np.int64)
for b in range(2):
    full[b * pb:(b + 1) * pb, b * pb:(b + 1) * pb] = blocks[b]; gfull[b * pb:(b + 1) * pb] = groups[b] + b * (pb // 5)
USEV = os.environ.get("USEV", "0") == "1"
Vs = [go.get_PCA_vectors(blocks[b], list(groups[b])) for b in range(2)]
Vfull = np.zeros((n, Vs[0].shape[1] + Vs[1].shape[1]), order="F")
Vfull[:pb, :Vs[0].shape[1]] = Vs[0]; Vfull[pb:, Vs[0].shape[1]:] = Vs[1]
Sf, _, objf, inf = kb.solve_s_group(full, gfull, "mvr", V=Vfull if USEV else None, ctx=kb.Context(0), return_info=True)
print("steps", inf["steps"], "accepted", inf["accepted"], "ndir", inf["n_directions"])
print("full: obj_init", inf["obj_init"], "obj", objf, "sweeps", inf["inner_sweeps"], "recomputed", go.group_block_objective(full, Sf, list(gfull), 1, "mvr"))
print("  trace", [(k, round(o, 4), round(d, 6)) for k, o, d in inf["trace"]])
slots = [None, None]; bar = threading.Barrier(2); out = [None, None]
def worker(rank):
    def reduce(op, vals):
        slots[rank] = vals.copy(); bar.wait(timeout=60)
        both = np.stack(slots); res = both.min(0) if op == 0 else (both.sum(0) if op == 1 else both.max(0))
        bar.wait(timeout=60); vals[:] = res
    out[rank] = kb.solve_s_group_sharded(blocks[rank], groups[rank], "mvr", reduce, V=Vs[rank] if USEV else None, ctx=kb.Context(0))
th = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(2)]
[t.start() for t in th]; [t.join(100) for t in th]
for r in range(2):
    S, obj, info = out[r]
    print("steps", info["steps"], "accepted", info["accepted"], "ndir", info["n_directions"])
    print("shard", r, "obj_init", info["obj_init"], "obj", obj, "sweeps", info["inner_sweeps"], "max|dS|", np.max(np.abs(S - Sf[r * pb:(r + 1) * pb, r * pb:(r + 1) * pb])))
    print("  trace", [(k, round(o, 4), round(d, 6)) for k, o, d in info["trace"]])
print("sum", out[0][1] + out[1][1])
