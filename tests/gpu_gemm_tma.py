"""B200: the TMA + mbarrier variant of the FP64 DMMA GEMM (csrc/dgemm_tma.cu: cp.async.bulk.tensor, 128-byte swizzle, one
producer warp) against the LDGSTS kernel the library uses (dgemm_dmma_kernel) and cuBLAS, same shapes, device pointers,
CUDA events, best of 5.  First a correctness pass on ragged shapes.  Prints a markdown table."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
d = torch.device("cuda", 0)
ctx = kb.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=d)
f64 = torch.float64


def run(which, M, N, K, A, B, C, lda, ldb, ldc, alpha=1.0, beta=0.0):
    if which == "tma":
        rc = ctx._lib.kb_test_dgemm_tma_dev(ctx.h, M, N, K, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc)
    else:
        rc = ctx._lib.kb_test_dgemm_dev(ctx.h, 0, 0, M, N, K, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc)
    kb.api.check(ctx.h, rc)


ok = True
for (M, N, K) in [(64, 64, 16), (64, 64, 64), (130, 70, 50), (257, 131, 77), (1000, 500, 333), (64, 64, 5000), (2048, 64, 64),
                  (64, 5000, 64), (1024, 1024, 1024), (4096, 4096, 512), (4096, 5000, 5000)]:
    lda, ldb, ldc = (M + 1) & ~1, (K + 1) & ~1, M
    A = torch.randn((K, lda), dtype=f64, device=d)
    B = torch.randn((N, ldb), dtype=f64, device=d)
    C0 = torch.randn((N, ldc), dtype=f64, device=d)
    C = C0.clone()
    torch.cuda.synchronize()             # the inputs are produced on torch's stream, the kernel runs on the library's
    run("tma", M, N, K, A, B, C, lda, ldb, ldc, alpha=0.7, beta=-1.3)
    ext.synchronize()
    want = 0.7 * (B[:, :K] @ A[:, :M]) - 1.3 * C0
    err = float((C - want).abs().max() / want.abs().max())
    print(f"check M={M} N={N} K={K}: max rel err {err:.2e}")
    if err >= 1e-12:
        E = ((C - want).abs() / want.abs().max())[:, :M]          # (N, M): E[n, m]
        nt, mt = (N + 63) // 64, (M + 63) // 64
        bad = [(mi, ni) for ni in range(nt) for mi in range(mt) if float(E[ni * 64:(ni + 1) * 64, mi * 64:(mi + 1) * 64].max()) >= 1e-12]
        print(f"   bad 64x64 tiles (m-tile, n-tile): {len(bad)} of {nt * mt}; first {bad[:12]} last {bad[-6:]}")
        mi, ni = bad[0]
        sub = E[ni * 64:(ni + 1) * 64, mi * 64:(mi + 1) * 64]
        print("   inside the first bad tile: bad rows(n)", sorted(set(torch.nonzero(sub >= 1e-12)[:, 0].tolist()))[:20],
              "bad cols(m)", sorted(set(torch.nonzero(sub >= 1e-12)[:, 1].tolist()))[:20])
    ok = ok and err < 1e-12
print("correctness:", "OK" if ok else "FAILED")
print()
print("| M | N | K | dgemm_dmma (LDGSTS) TFLOP/s | dgemm_tma (UTMALDG) TFLOP/s | cuBLAS TFLOP/s | tma / ldgsts | tma / cuBLAS |")
print("|---|---|---|---|---|---|---|---|")
for (M, N, K) in [(8192, 8192, 8192), (4096, 5000, 5000), (4096, 5000, 10000), (16384, 5000, 5000), (4096, 5056, 5056)]:
    A = torch.randn((K, M), dtype=f64, device=d)
    B = torch.randn((N, K), dtype=f64, device=d)
    C = torch.empty((N, M), dtype=f64, device=d)
    torch.cuda.synchronize()
    best = {"ldgsts": 1e9, "tma": 1e9, "cublas": 1e9}
    for it in range(5):
        for which in ("ldgsts", "tma"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext); run(which, M, N, K, A, B, C, M, K, M); e1.record(ext); ext.synchronize()
            best[which] = min(best[which], e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); R = torch.matmul(B, A); e1.record(); torch.cuda.synchronize()
        best["cublas"] = min(best["cublas"], e0.elapsed_time(e1))
    fl = 2.0 * M * N * K / 1e9
    print(f"| {M} | {N} | {K} | {fl / best['ldgsts']:.2f} | {fl / best['tma']:.2f} | {fl / best['cublas']:.2f} | "
          f"{best['ldgsts'] / best['tma']:.3f} | {best['cublas'] / best['tma']:.3f} |")
