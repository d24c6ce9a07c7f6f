"""B200 profiling driver: ONLY the sampling stage (kb_sample_dev) on factor blocks of the diagonal-S form, so that ncu can
pick the batched copy GEMM by launch index.  dgemm_dmma_kernel launches: per 4096-row chunk [mean GEMM, batched copies]."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
n, p, m = int(os.environ.get("N", 8192)), int(os.environ.get("P", 5000)), int(os.environ.get("M", 5))
d = torch.device("cuda", 0)
ctx = kb.Context(0)
g = torch.Generator(device=d); g.manual_seed(1)
f64 = torch.float64
Xt = torch.randn((p, n), dtype=f64, device=d, generator=g)
Sigma = torch.from_numpy(np.ascontiguousarray(kb.harness.block_sigma(p, 10, 0.5, 0.25))).to(d)
s = torch.full((p,), 0.3, dtype=f64, device=d)
SiS = torch.empty((p, p), dtype=f64, device=d)
Lb = torch.empty((2 * m - 1, p, p), dtype=f64, device=d)
kb.dev.cond_factor_dev(ctx, Sigma.data_ptr(), p, s.data_ptr(), True, m, SiS.data_ptr(), Lb.data_ptr())
mu = torch.zeros(p, dtype=f64, device=d)
Xko = torch.empty((m * p, n), dtype=f64, device=d)
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
for it in range(int(os.environ.get("REPS", 1))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    last = it == int(os.environ.get("REPS", 1)) - 1
    if last:
        torch.cuda.nvtx.range_push("prof")          # ncu --nvtx --nvtx-include "prof/" -k regex:dgemm_dmma -s 1 -c 1
    kb.dev.sample_dev(ctx, Xt.data_ptr(), n, n, p, mu.data_ptr(), SiS.data_ptr(), Lb.data_ptr(), m, Xko.data_ptr(), n, seed=7)
    if last:
        torch.cuda.nvtx.range_pop()
    e1.record(ext); ext.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"sample_dev n={n} p={p} m={m}: {ms:.2f} ms, {(2.0 * m + 1) * n * p * p / ms / 1e9:.2f} TFLOP/s executed ((2m+1) n p^2), "
          f"{3.0 * m * n * p * p / ms / 1e9:.2f} in SURVEY's 3 m n p^2 count")
