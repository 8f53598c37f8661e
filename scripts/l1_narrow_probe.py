"""Narrow DAG layer 1 (cfg4: 6300 rows, d = 63 -> K = 64, width 630) through the tensor-core GEMM engine against the saved
embedding plane: is it worth routing the backward GEMMs there?  (run on the GPU box)"""
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

M, N, K = 6300, 630, 64
E = torch.randn(M, K, device="cuda"); E[:, 63] = 0
W = torch.zeros(N, K, device="cuda"); W[:, :63] = torch.randn(N, 63, device="cuda") / 8
T = torch.randn(63, N, device="cuda")
dY = torch.zeros(M, 632, device="cuda")[:, :N]; dY.copy_(torch.randn(M, N, device="cuda"))
for mode in ("ffma", "tf32x3"):
    G.ops.set_gemm_mode(mode)
    y = torch.zeros(M, 632, device="cuda")
    f = t(lambda: G.ops.linear_fwd(E, W, T, relu=True, bias_period=63, out=y, ldy=632))
    w = t(lambda: G.ops.linear_wgrad(dY, 632, E, K, M, N, K))
    d = t(lambda: G.ops.linear_dgrad(dY, 632, W, None, M))
    print(f"{mode}: fwd {f:.1f} us  wgrad {w:.1f} us  dgrad {d:.1f} us")
    if mode == "tf32x3":
        ref = torch.relu(E.double() @ W.double().t() + T.double().repeat(M // 63, 1))
        print("   fwd err", float((y[:, :N].double() - ref).abs().max()),
              "wgrad err", float((G.ops.linear_wgrad(dY, 632, E, K, M, N, K).double() - dY.double().t() @ E.double()).abs().max()),
              "dgrad err", float((G.ops.linear_dgrad(dY, 632, W, None, M).double() - dY.double() @ W.double()).abs().max()))
G.ops.set_gemm_mode("ffma")
