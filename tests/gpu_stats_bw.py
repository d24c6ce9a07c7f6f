"""B200: achieved HBM GB/s of the marginal-statistics pass (kb_marginal_stats_dev) at BASELINE configs[2] size."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
n, p, m = int(os.environ.get("N", 100000)), int(os.environ.get("P", 5000)), int(os.environ.get("M", 5))
d = torch.device("cuda", 0)
ctx = kb.Context(0)
g = torch.Generator(device=d); g.manual_seed(1)
Xt = torch.randn((p, n), dtype=torch.float64, device=d, generator=g)
Xko_t = torch.randn((m * p, n), dtype=torch.float64, device=d, generator=g)
y = torch.randn(n, dtype=torch.float64, device=d, generator=g)
T = torch.empty((m + 1) * p, dtype=torch.float64, device=d)
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
best = 1e9
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    rc = ctx._lib.kb_marginal_stats_dev(ctx.h, y.data_ptr(), Xt.data_ptr(), n, Xko_t.data_ptr(), n, n, p, m, T.data_ptr())
    e1.record(ext); ext.synchronize()
    assert rc == 0
    best = min(best, e0.elapsed_time(e1))
bytes_ = 8.0 * n * (m + 1) * p
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
# spot check against torch
xs = (Xt[:3] - Xt[:3].mean(1, keepdim=True)) / Xt[:3].std(1, keepdim=True)
ys = (y - y.mean()) / y.std()
ref = (xs @ ys) ** 2 / n
print(json.dumps({"kernel": "marginal_accumulate_kernel", "n": n, "p": p, "m": m, "ms": best, "algorithmic_bytes": bytes_,
                  "achieved_gbs": bytes_ / best / 1e6, "peak_gbs": peak, "frac": bytes_ / best / 1e6 / peak,
                  "max_rel_err_vs_torch": float(((T[:3] - ref).abs() / ref.abs()).max())}))
