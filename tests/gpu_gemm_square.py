"""B200: dgemm_dmma_kernel against cuBLAS DGEMM on the SAME plain shapes (device pointers, CUDA events, best of 5)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
d = torch.device("cuda", 0)
ctx = kb.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
f64 = torch.float64
for (M, N, K) in [(8192, 8192, 8192), (4096, 5000, 5000), (4096, 5000, 10000), (16384, 5000, 5000)]:
    A = torch.randn((K, M), dtype=f64, device=d)       # column-major M x K  == row-major (K, M)
    B = torch.randn((N, K), dtype=f64, device=d)       # column-major K x N
    Cc = torch.empty((N, M), dtype=f64, device=d)      # column-major M x N
    torch.cuda.synchronize()
    best_o = best_c = 1e9
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        rc = ctx._lib.kb_test_dgemm_dev(ctx.h, 0, 0, M, N, K, 1.0, A.data_ptr(), M, B.data_ptr(), K, 0.0, Cc.data_ptr(), M)
        e1.record(ext); ext.synchronize()
        assert rc == 0
        best_o = min(best_o, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); R = torch.matmul(B, A); e1.record(); torch.cuda.synchronize()     # (N,K)@(K,M) = C' row-major = C col-major
        best_c = min(best_c, e0.elapsed_time(e1))
    err = float((R - Cc).abs().max() / R.abs().max())
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: dgemm_dmma_kernel {fl / best_o / 1e9:6.2f} TFLOP/s, cuBLAS {fl / best_c / 1e9:6.2f} TFLOP/s, "
          f"ratio {best_c / best_o:.3f}, max rel diff {err:.1e}")
