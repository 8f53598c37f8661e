import os
"""Warp-role timeline of the resident wgrad kernel (CTA 0); run on the GPU box.  usage: wg_trace.py M N K passes"""
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
NPn, KP = (N + 31) // 32 * 32, (K + 31) // 32 * 32
X = torch.zeros(M, KP, device="cuda"); X[:, :K] = torch.randn(M, K, device="cuda")
dY = torch.zeros(M, NPn, device="cuda"); dY[:, :N] = torch.randn(M, N, device="cuda")
fn = lambda: G.ops.linear_wgrad_rw(dY, X, N, K, passes=passes)
for _ in range(3): fn()
buf = torch.zeros(3 * 256, dtype=torch.int64, device="cuda")
lib.gnf_linear_wgrad_rw_set_trace(C.c_void_p(buf.data_ptr()))
fn(); torch.cuda.synchronize()
lib.gnf_linear_wgrad_rw_set_trace(None)
t = buf.cpu().view(3, 256)
t0 = int(t[t > 0].min())
print(f"M={M} N={N} K={K} passes={passes}: SM clocks relative to the first stamp")
for r, n in enumerate(["issuer (per chunk: start, A full, image full, issued)", "stager (per chunk: start, raw landed, image free, published)",
                       "loader (per chunk: start, split + A free, handed over)"]):
    v = [int(x) - t0 for x in t[r] if int(x) > 0]
    print(n)
    per = 3 if r == 2 else 4
    for i in range(0, min(len(v), 8 * per), per):
        print("   ", v[i:i + per])
