"""Multi-GPU through the C ABI (kb_create_multi) on all visible GPUs: row-sharded Gram with the peer-to-peer all-reduce
kernel, row-sharded one-shot knockoffs, :mvr || :sdp_ccd replicas, LD blocks of configs[3].  Prints one JSON object.
Run with KB_MULTI_TRACE=1 to get the per-device all-reduce kernel times on stderr."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
nd = torch.cuda.device_count()
devs = list(range(nd)) if nd > 1 else [0, 0]
out = {"devices": devs}
p = int(os.environ.get("P", 5000)); n = int(os.environ.get("N", 100000)); m = int(os.environ.get("M", 5))
Sigma = kb.harness.block_sigma(p, 10, 0.5, 0.25)
rng = np.random.default_rng(1)
X = np.asfortranarray(rng.standard_normal((n, p)))          # pageable host memory, like a Julia Matrix{Float64}
one = kb.Context(0)
mc = kb.MultiContext(devs)
for name, fn in (("gram_1gpu", lambda: kb.gram(X, ctx=one)), ("gram_multi", lambda: mc.gram(X))):
    fn(); t = time.perf_counter(); G, mu = fn(); out[name + "_s"] = time.perf_counter() - t
    out[name + "_checksum"] = float(np.sum(G))
mu = np.zeros(p)
for name, fn in (("modelx_1gpu", lambda: kb.modelX_gaussian_knockoffs(X, "mvr", mu, Sigma, m=m, seed=3, ctx=one)),
                 ("modelx_multi", lambda: mc.modelX_gaussian_knockoffs(X, "mvr", mu, Sigma, m=m, seed=3))):
    r = fn(); t = time.perf_counter(); r = fn(); dt = time.perf_counter() - t
    out[name + "_s"] = dt; out[name + "_rows_per_s"] = n / dt; out[name + "_checksum"] = float(np.sum(r.Xko[::97, ::89]))
    del r
t = time.perf_counter(); ss, infos = mc.solve_s_many(Sigma, ["mvr", "sdp_ccd"], [m, m]); out["solve_many_mvr_sdp_s"] = time.perf_counter() - t
out["solve_many_ms"] = [i["solve_ms"] + i["init_ms"] for i in infos]
blocks, groups = [], []
for b in range(8):
    g = np.repeat(np.arange(1, 251), 5); rho = 0.7 + 0.02 * b
    S = np.where(g[:, None] == g[None, :], rho, rho * 0.25).astype(float); np.fill_diagonal(S, 1.0)
    blocks.append(np.asfortranarray(S)); groups.append(g)
mc.solve_s_group_blockdiag([b[:50, :50] for b in blocks], [g[:50] for g in groups], "mvr")
t = time.perf_counter(); Sb, obj, gi = mc.solve_s_group_blockdiag(blocks, groups, "mvr"); out["config4_group_mvr_8blocks_s"] = time.perf_counter() - t
out["config4_obj"] = obj; out["config4_inner_sweeps"] = gi[0]["inner_sweeps"]
print(json.dumps(out))
