"""Diagnostics for the tcgen05 conventions (run on the GPU box): prints where the self-test GEMM deviates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G

torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
for mode in (1, 0):
    for (N, K) in [(16, 8), (32, 16), (160, 152)]:
        r = torch.arange(1, 129, device="cuda").float().view(-1, 1)
        n = torch.arange(1, N + 1, device="cuda").float().view(-1, 1)
        ok_all = True
        for k0 in sorted({0, 1, 3, 4, 7, K - 1}):
            A = torch.zeros(128, K, device="cuda"); A[:, k0] = r[:, 0]
            W = torch.zeros(N, K, device="cuda"); W[:, k0] = n[:, 0]
            try:
                Cm = G.ops.tc_selftest(A, W, mode)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"mode={mode} N={N} K={K} k0={k0}: EXCEPTION {e}")
                ok_all = False
                break
            ref = r @ n.t()
            if not torch.equal(Cm, ref):
                ok_all = False
                bad = (Cm != ref)
                print(f"mode={mode} N={N} K={K} k0={k0}: MISMATCH {int(bad.sum())}/{bad.numel()} entries; rows bad: "
                      f"{bad.any(1).nonzero().flatten()[:12].tolist()} cols bad: {bad.any(0).nonzero().flatten()[:12].tolist()}")
                print(" got [0:4,0:8]\n", Cm[0:4, 0:8].cpu(), "\n want\n", ref[0:4, 0:8].cpu())
        print(f"mode={mode} N={N} K={K}: {'OK' if ok_all else 'FAIL'}")
