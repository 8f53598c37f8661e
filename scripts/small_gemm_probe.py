"""The skinny layers of the cfg4 step (conditioner output layer 630 -> 30, the conditioning half of the integrand's first layer
30 -> 150) on the FFMA engine vs the tensor-core engine (v1, planned tiles).  (run on the GPU box)"""
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

def pad(M, N):
    return torch.zeros(M, (N + 3) // 4 * 4, device="cuda")[:, :N]

for (M, N, K) in [(6300, 30, 630), (6300, 150, 30), (6300, 630, 64)]:
    X = pad(M, K); X.copy_(torch.randn(M, K, device="cuda").relu())
    W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
    dY = pad(M, N); dY.copy_(torch.randn(M, N, device="cuda"))
    for mode in ("ffma", "tf32x3"):
        G.ops.set_gemm_mode(mode)
        try:
            f = t(lambda: G.ops.linear_fwd(X, W, b, relu=False))
            d = t(lambda: G.ops.linear_dgrad(dY, dY.stride(0), W, X, M))
            w = t(lambda: G.ops.linear_wgrad(dY, dY.stride(0), X, X.stride(0), M, N, K))
            c = t(lambda: G.ops.colsum(dY, dY.stride(0), M, N))
            print(f"M={M} N={N} K={K} {mode}: fwd {f:.1f} us  dgrad(mask) {d:.1f} us  wgrad {w:.1f} us  colsum {c:.1f} us")
        except Exception as e:
            print(f"M={M} N={N} K={K} {mode}: {e}")
G.ops.set_gemm_mode("ffma")
