"""Layer 1 of the DAG conditioner at a wide-flow shape: forward output, dW1, dA, dx of the embedding-plane / tensor-core path and
of the FFMA loader kernels against a float64 torch evaluation with the same (dumped) gate noise.  usage: dag_l1_accuracy.py [B d N1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnf_b200 as G  # noqa: E402
from gnf_b200 import ops  # noqa: E402


def l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main(B=16, d=784, N1=1024):
    dev = "cuda"
    torch.manual_seed(0)
    T = 1.0
    gate = ops.GateSpec(G._lib.GATE_GUMBEL, G._lib.IMP_SOFT, 0., T, seed=5, offset=3)
    x0 = torch.randn(B, d, device=dev)
    A0 = (torch.rand(d, d, device=dev) < .05).float() * torch.rand(d, d, device=dev) * 2      # sparse-ish importance, like the MNIST prior
    W0 = torch.randn(N1, 2 * d, device=dev) / d ** .5
    b0 = torch.randn(N1, device=dev) * .1
    gh = torch.randn(B, d, N1, device=dev)
    n1, n2 = ops.dag_dump_noise(gate, B, d, dev)
    # float64 reference (DAGConditioner.py:94-103, 118-153)
    x, A, W, b = x0.double().requires_grad_(), A0.double().requires_grad_(), W0.double().requires_grad_(), b0.double().requires_grad_()
    P = 2 * (torch.sigmoid(2 * A ** 2) - .5)
    eps = 1e-6
    g1, g2 = -torch.log(-torch.log(n1.double())), -torch.log(-torch.log(n2.double()))
    z1 = torch.exp((torch.log(P + eps) + g1) / T)
    z2 = torch.exp((torch.log(1 - P + eps) + g2) / T)
    Gt = z1 / (z1 + z2)                                                   # [B, d, d]
    e = x.unsqueeze(1) * Gt
    y = e @ W[:, :d].t() + W[:, d:].t().unsqueeze(0) + b
    y.backward(gh.double())
    ref = [y.detach(), x.grad, A.grad, W.grad, b.grad]
    print(f"B={B} d={d} N1={N1}: relative L2 against float64")
    for name, mode, plane in (("plane + 3xTF32", "tf32x3", True), ("FFMA loader kernels", "ffma", False)):
        ops.set_gemm_mode(mode)
        ops.DAG_L1_PLANE = plane
        xg, Ag, Wg, bg = x0.clone().requires_grad_(), A0.clone().requires_grad_(), W0.clone().requires_grad_(), b0.clone().requires_grad_()
        h = ops.DagMlpFn.apply(xg, Ag, gate, True, Wg, bg)
        h.backward(gh)
        out = [h.detach(), xg.grad, Ag.grad, Wg.grad, bg.grad]
        print(f"  {name:22s} " + "  ".join(f"{n} {l2(a, r):.2e}" for n, a, r in zip(["y", "dx", "dA", "dW1", "db1"], out, ref)) +
              f"   mean signed y error / mean |y| {float((out[0].double() - ref[0]).mean() / ref[0].abs().mean()):+.2e}")
    ops.DAG_L1_PLANE = True


if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
