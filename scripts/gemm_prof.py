"""Target for ncu: three launches of one GEMM shape.  usage: gemm_prof.py mode M N K op"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (20000, 1024, 1024)
op = sys.argv[5] if len(sys.argv) > 5 else "fwd"
X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
dY = torch.randn(M, N, device="cuda")
G.ops.set_gemm_mode(mode)
fn = {"fwd": lambda: G.ops.linear_fwd(X, W, b, relu=True), "dgrad": lambda: G.ops.linear_dgrad(dY, N, W, X, M),
      "wgrad": lambda: G.ops.linear_wgrad(dY, N, X, K, M, N, K)}[op]
for _ in range(3):
    fn()
torch.cuda.synchronize()
