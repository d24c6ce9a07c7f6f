"""B200: time of one p x p Cholesky factorisation (kb_test_potrf is host-pointer; this uses cond_factor m = 1 pieces
through the device entry instead): look-ahead on / off via KB_POTRF_LOOKAHEAD."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
p = int(os.environ.get("P", 5000))
ctx = kb.Context(0)
d = torch.device("cuda", 0)
Sigma = torch.from_numpy(np.ascontiguousarray(kb.harness.block_sigma(p, 10, 0.5, 0.25))).to(d)
s = torch.full((p,), 0.3, dtype=torch.float64, device=d)
SiS = torch.empty((p, p), dtype=torch.float64, device=d)
for m in (1, 5):
    Lb = torch.empty((2 * m - 1, p, p), dtype=torch.float64, device=d)
    ext = torch.cuda.ExternalStream(ctx.stream, device=d)
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        kb.dev.cond_factor_dev(ctx, Sigma.data_ptr(), p, s.data_ptr(), True, m, SiS.data_ptr(), Lb.data_ptr())
        e1.record(ext); ext.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cond_factor p={p} m={m}: {best:.2f} ms  (lookahead={os.environ.get('KB_POTRF_LOOKAHEAD', '1')})")
t = time.perf_counter(); lam = kb.eigmin(Sigma.cpu().numpy(), ctx=ctx); print(f"eigmin {lam:.15f} in {1e3 * (time.perf_counter() - t):.1f} ms (incl. H2D)")
