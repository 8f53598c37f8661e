import os
"""Ablation timing of the resident-weight GEMM: which of loads / stores / MMAs bounds it (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
lib = G._lib.lib()
M, N, K = 1419264, 150, 150
X = torch.zeros(M, 160, device="cuda"); X[:, :K] = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for bits, name in ((0, "full"), (1, "no stores"), (2, "no loads"), (4, "no MMAs"), (3, "no loads, no stores"), (5, "no stores, no MMAs"),
                   (6, "no loads, no MMAs"), (7, "nothing but the hand-shakes")):
    lib.gnf_linear_rw_set_debug(bits)
    r = []
    for passes in (1, 3):
        r.append(t(lambda: G.ops.linear_fwd_rw(X, W, b, relu=True, passes=passes, want_bits=True)))
    print(f"{name:32s} passes=1 {r[0]:6.1f} us   passes=3 {r[1]:6.1f} us")
lib.gnf_linear_rw_set_debug(0)
