import os
"""Tile-width / split-K sweep of the generic tcgen05 GEMM engine (run on the GPU box): forced tiles vs the planner."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()

lib = G._lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def t(fn, n=12):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

for (M, N, K) in [(6300, 632, 632), (64512, 632, 632), (10000, 212, 212), (78400, 1024, 1024)]:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
    dY = torch.randn(M, N, device="cuda")
    for mode, bns in (("tf32x3", (0, 96, 128, 160)), ("tf32", (0, 128, 160, 192, 224, 256))):
        G.ops.set_gemm_mode(mode)
        row = f"M={M} N={N} K={K} {mode}: "
        for bn in bns:
            lib.gnf_tc_gemm_set_tile(bn, 0)
            a = t(lambda: G.ops.linear_fwd(X, W, b, relu=True))
            c = t(lambda: G.ops.linear_dgrad(dY, N, W, X, M))
            row += f" bn={bn}: fwd {a:.0f} dgrad {c:.0f} |"
        print(row, flush=True)
    G.ops.set_gemm_mode("tf32x3")
    row = f"M={M} N={N} K={K} wgrad 3x: "
    for bn in (0, 128, 160):
        for sp in ((0,) if bn == 0 else (4, 5, 6, 7, 8, 12, 16, 22)):
            lib.gnf_tc_gemm_set_tile(bn, sp)
            d = t(lambda: G.ops.linear_wgrad(dY, N, X, K, M, N, K))
            row += f" bn={bn},sp={sp}: {d:.0f} |"
    print(row, flush=True)
    lib.gnf_tc_gemm_set_tile(0, 0)
