"""B200: device-resident timings of the other BASELINE.json configs (the bench line is configs[2]).
c1 AR(1) p=500 n=1000 :mvr m=1 | c2 AR(1) p=2000 n=10000 :maxent m=1 | c4 one LD block of the group config (p=1250, groups of 5,
:mvr) | c5 dosage-like X n=500000 p=5000: Gram + :maxent + sampling m=1.   Prints one JSON object."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
d = torch.device("cuda", 0)
ctx = kb.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
f64 = torch.float64
out = {}


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext); r = fn(); e1.record(ext); ext.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


def pipeline(Sigma_t, Xt, method, m):
    p, n = Xt.shape
    s = torch.empty(p, dtype=f64, device=d); SiS = torch.empty((p, p), dtype=f64, device=d)
    Lb = torch.empty((2 * m - 1, p, p), dtype=f64, device=d); Xko = torch.empty((m * p, n), dtype=f64, device=d)
    mu = torch.zeros(p, dtype=f64, device=d)
    t_s, info = timed(lambda: kb.dev.solve_s_dev(ctx, Sigma_t.data_ptr(), p, method, m, s.data_ptr()))
    t_f, _ = timed(lambda: kb.dev.cond_factor_dev(ctx, Sigma_t.data_ptr(), p, s.data_ptr(), True, m, SiS.data_ptr(), Lb.data_ptr()))
    t_x, _ = timed(lambda: kb.dev.sample_dev(ctx, Xt.data_ptr(), n, n, p, mu.data_ptr(), SiS.data_ptr(), Lb.data_ptr(), m, Xko.data_ptr(), n, seed=1))
    return {"solve_s_ms": t_s, "sweeps": info["sweeps"], "cond_factor_ms": t_f, "sampling_ms": t_x, "rows_per_s": n / ((t_s + t_f + t_x) * 1e-3)}


g = torch.Generator(device=d); g.manual_seed(20260001)
for name, p, n, method in (("c1_ar1_p500_n1000_mvr", 500, 1000, "mvr"), ("c2_ar1_p2000_n10000_maxent", 2000, 10000, "maxent")):
    S = torch.from_numpy(np.ascontiguousarray(kb.harness.ar1_sigma(p, 0.5))).to(d)
    Xt = torch.linalg.cholesky(S) @ torch.randn((p, n), dtype=f64, device=d, generator=g)
    out[name] = pipeline(S, Xt.contiguous(), method, 1)

# c4: one LD block of the group configuration (the 8 blocks are independent; dist.solve_s_group_blockdiag shards them)
pb = 1250
grp = np.repeat(np.arange(1, pb // 5 + 1), 5)
Sg = np.where(grp[:, None] == grp[None, :], 0.75, 0.75 * 0.25).astype(float); np.fill_diagonal(Sg, 1.0)
kb.solve_s_group(np.asfortranarray(Sg[:50, :50]), grp[:50], "mvr", ctx=ctx)
t = time.perf_counter(); S_, _, obj, info = kb.solve_s_group(np.asfortranarray(Sg), grp, "mvr", ctx=ctx, return_info=True); wall = time.perf_counter() - t
out["c4_group_mvr_one_block_p1250"] = {"wall_s": wall, "solve_ms": info["solve_ms"], "inner_sweeps": info["inner_sweeps"], "steps": info["steps"], "obj": obj}

# c5: dosage-like X generated on the device (0/1/2, allele frequencies U(0.05, 0.5), LD within blocks of 50 through a shared latent)
n, p = int(os.environ.get("C5_ROWS", 500000)), 5000
Xt = torch.empty((p, n), dtype=f64, device=d)
freq = 0.05 + 0.45 * torch.rand(p, dtype=f64, device=d, generator=g)
thr = torch.special.ndtri(freq)[:, None]
for b0 in range(0, p, 50):
    b1 = min(p, b0 + 50)
    acc = torch.zeros((b1 - b0, n), dtype=f64, device=d)
    for _ in range(2):
        z = torch.randn((1, n), dtype=f64, device=d, generator=g)
        e = torch.randn((b1 - b0, n), dtype=f64, device=d, generator=g)
        acc += ((0.6 ** 0.5) * z + (0.4 ** 0.5) * e < thr[b0:b1]).to(f64)
    Xt[b0:b1] = acc
del acc, z, e
torch.cuda.synchronize()
t_g, (Sig, mu) = timed(lambda: kb.dist.gram_sharded_dev(ctx, Xt, n), reps=2)
res = pipeline(Sig, Xt, "maxent", 1)
res.update({"gram_ms": t_g, "gram_tflops": n * p * (p + 1.0) / t_g / 1e9, "n": n, "p": p})
res["rows_per_s_incl_gram"] = n / ((t_g + res["solve_s_ms"] + res["cond_factor_ms"] + res["sampling_ms"]) * 1e-3)
out["c5_dosage_n500000_p5000_maxent"] = res
print(json.dumps(out))
