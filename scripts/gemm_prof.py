import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G
M, N, K = 20000, 1024, 1024
X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** .5; b = torch.randn(N, device="cuda")
G.ops.set_gemm_mode(sys.argv[1] if len(sys.argv) > 1 else "tf32")
for _ in range(3):
    G.ops.linear_fwd(X, W, b, relu=True)
torch.cuda.synchronize()
