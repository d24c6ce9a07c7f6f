"""B200: device time of solve_equi (= the eigmin probes) and of the whole :mvr solve at p = 5000 (device pointers)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
p = int(os.environ.get("P", 5000))
ctx = kb.Context(0)
d = torch.device("cuda", 0)
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
for name, Sig in (("block", kb.harness.block_sigma(p, 10, 0.5, 0.25)), ("ar1", kb.harness.ar1_sigma(p, 0.5))):
    Sigma = torch.from_numpy(np.ascontiguousarray(Sig)).to(d)
    s = torch.empty(p, dtype=torch.float64, device=d)
    for method in ("equi", "mvr"):
        best = 1e9
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            info = kb.dev.solve_s_dev(ctx, Sigma.data_ptr(), p, method, 5, s.data_ptr())
            e1.record(ext); ext.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name} p={p} {method}: {best:.2f} ms  launches={info['launches']} sweeps={info['sweeps']} ccd={info['solve_ms']:.1f} init={info['init_ms']:.1f}  s[0]={float(s[0]):.15f}")
