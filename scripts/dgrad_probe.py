"""Where do the 15 us between the engine-v2 forward (40 us) and dgrad (55 us) at 6300 x 632 x 632 go: the MN-major weight operand or
the ReLU-mask read of the epilogue?  (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G

def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

G.ops.set_gemm_mode("tf32x3")
for (M, N, K) in [(6300, 632, 632), (64512, 632, 632), (78400, 1024, 1024)]:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
    dY = torch.randn(M, N, device="cuda"); WT = W.t().contiguous()
    print(f"M={M} N={N} K={K}: fwd bias+relu {t(lambda: G.ops.linear_fwd(X, W, b, relu=True)):.1f} us | "
          f"fwd plain {t(lambda: G.ops.linear_fwd(X, W, None, relu=False)):.1f} | "
          f"dgrad MN-major W + mask {t(lambda: G.ops.linear_dgrad(dY, N, W, X, M)):.1f} | "
          f"dgrad MN-major W, no mask {t(lambda: G.ops.linear_dgrad(dY, N, W, None, M)):.1f} | "
          f"dgrad as forward on W^T (K-major), no mask {t(lambda: G.ops.linear_fwd(dY, WT, None, relu=False)):.1f}")
