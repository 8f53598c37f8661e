import os
"""Warp-role timeline of the resident-weight layer GEMM (CTA 0); run on the GPU box.  usage: rw_trace.py M N K passes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
lib = G._lib.lib()
M, N, K, passes = (int(v) for v in sys.argv[1:5])
NP, KP = (N + 31) // 32 * 32, (K + 31) // 32 * 32   # X is padded to whole 32-column chunks
X = torch.zeros(M, KP, device="cuda"); X[:, :K] = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
fn = lambda: G.ops.linear_fwd_rw(X, W, b, relu=True, passes=passes, want_bits=True)
for _ in range(3): fn()
buf = torch.zeros(4 * 256, dtype=torch.int64, device="cuda")
lib.gnf_linear_rw_set_trace(C.c_void_p(buf.data_ptr()))
fn(); torch.cuda.synchronize()
lib.gnf_linear_rw_set_trace(None)
t = buf.cpu().view(4, 256)
t0 = int(t[t > 0].min())
nch = KP // 32
print(f"M={M} N={N} K={K} passes={passes}: SM clocks relative to the first stamp")
names = [f"mma (per tile: start, D free, {nch} x chunk issued, all issued)", "loader even chunks (per chunk: start, half staged + A free, handed over)",
         "loader odd chunks", f"epilogue (per tile: start, D full, first block, {NP // 32} x block stored)"]
per = [3 + nch, 3, 3, 3 + NP // 32]
for r, n in enumerate(names):
    v = [int(x) - t0 for x in t[r] if int(x) > 0]
    print(n)
    for i in range(0, min(len(v), per[r] * (10 if r in (1, 2) else 5)), per[r]):
        print("   ", v[i:i + per[r]])
