"""Skinny layers (csrc/thin.cuh) against the register-tiled GEMM on the shapes the hot path has: CUDA-event time per call, dev build.
usage: thin_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402

lib = devlib.install()
import gnf_b200 as G  # noqa: E402


def timed(fn, n=20):
    """GPU time per call: n calls captured in a CUDA graph (the Python wrapper costs ~20 us per call, more than these kernels)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1000


G.ops.set_gemm_mode("ffma")
# (rows, out features, in features, what)
cases = [(6300, 30, 630, "cfg4 conditioner output layer"), (6300, 150, 30, "cfg4 integrand first layer, conditioning half"),
         (78400, 2, 1024, "cfg5 conditioner output layer"), (10000, 30, 210, "cfg3-like output layer")]
for M, N, K, what in cases:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
    dY = torch.randn(M, N, device="cuda")
    row = []
    for thin in (1, 0):
        lib.gnf_linear_set_thin(thin)
        f = timed(lambda: G.ops.linear_fwd(X, W, b, relu=False))
        d = timed(lambda: G.ops.linear_dgrad(dY, N, W, X, M))
        row.append((f, d))
    print(f"{what:48s} M={M:6d} N={N:4d} K={K:4d}: forward {row[0][0]:6.1f} us (tile GEMM {row[1][0]:6.1f})   dgrad {row[0][1]:6.1f} us (tile GEMM {row[1][1]:6.1f})")
lib.gnf_linear_set_thin(1)
