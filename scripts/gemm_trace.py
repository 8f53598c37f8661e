import os
"""Warp-role timeline of the tcgen05 GEMM engine (CTA 0) for one shape; run on the GPU box.
usage: gemm_trace.py M N K op(fwd|dgrad|wgrad) passes(1|3)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
lib = G._lib.lib()
M, N, K = (int(v) for v in sys.argv[1:4])
op, passes = sys.argv[4], int(sys.argv[5])
G.ops.set_gemm_mode("tf32" if passes == 1 else "tf32x3")
# activations as the model allocates them: row stride padded to 4 floats (TMA-loadable)
X = G.ops._rows(M, K, torch.empty(1, device="cuda")); X.normal_()
W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
dY = G.ops._rows(M, N, torch.empty(1, device="cuda")); dY.normal_()
fn = {"fwd": lambda: G.ops.linear_fwd(X, W, b, relu=True), "dgrad": lambda: G.ops.linear_dgrad(dY, dY.stride(0), W, X, M),
      "wgrad": lambda: G.ops.linear_wgrad(dY, dY.stride(0), X, X.stride(0), M, N, K)}[op]
if len(sys.argv) > 6:
    lib.gnf_tc_gemm_set_v2(int(sys.argv[6]))      # dev build: 0 = first engine for the pre-split forward / dgrad too
for _ in range(3): fn()
buf = torch.zeros(8 * 256, dtype=torch.int64, device="cuda")
lib.gnf_tc_gemm_set_trace(C.c_void_p(buf.data_ptr()))
fn(); torch.cuda.synchronize()
lib.gnf_tc_gemm_set_trace(None)
t = buf.cpu().view(8, 256)
t0 = int(t[t > 0].min())
names = ["tma issued", "stager landed", "stager published", "mma chunk ready", "mma tile committed", "epi start", "epi end",
         "epi tile1 chunk phases (start, acc+aux issued, computed, staged, stored, next acc ready) x chunks | engine v2: output pass per tile and 32-column block (start, staged, stored)"]
print(f"M={M} N={N} K={K} {op} passes={passes}: SM clocks relative to the first stamp (engine v2 rows: tma issued, A-writer landed, A-writer published, "
      f"mma chunk ready, mma group committed, fold start, fold end)")
for r, n in enumerate(names):
    v = [int(x) - t0 for x in t[r] if int(x) > 0][:(66 if r == 7 else 24)]
    print(f"{n:20s}", v)
