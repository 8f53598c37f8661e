"""Fused strict UMNN kernels (gnf_umnn_fwd_tc3 / gnf_umnn_bwd_tc3): plain launch against clusters of two CTAs that share one
multicast read of every streamed weight chunk.  Same inputs through the public op (UmnnFn), outputs / gradients compared, CUDA
event timings of the forward and of the backward.  Dev build.  usage: u3_cluster.py [B d S I E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402

lib = devlib.install()
import gnf_b200 as G  # noqa: E402
from gnf_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main(B=100, d=63, S=20, I=150, E=30, nh=3):
    dev = "cuda"
    torch.manual_seed(0)
    ops.set_gemm_mode("auto")
    dims = [1 + E] + [I] * nh + [1]
    Ws = [(torch.empty(dims[l + 1], dims[l], device=dev).uniform_(-1, 1) / dims[l] ** 0.5).requires_grad_() for l in range(len(dims) - 1)]
    bs = [(torch.empty(dims[l + 1], device=dev).uniform_(-1, 1) / dims[l] ** 0.5).requires_grad_() for l in range(len(dims) - 1)]
    params = [t for pair in zip(Ws, bs) for t in pair]
    x = torch.randn(B, d, device=dev, requires_grad=True)
    h = (torch.randn(B, d, E, device=dev) * 0.5).requires_grad_()
    gz = torch.randn(B, d, device=dev)
    gl = torch.randn(B, device=dev)
    res = {}
    for cs in (1, 2, 1, 2):
        lib.gnf_umnn_tc3_set_cluster(cs)
        for it in range(14):
            if it == 4:
                ops.enable_kernel_timing(True)
            for t in [x, h] + params:
                t.grad = None
            z, jac, logdet, _ = ops.UmnnFn.apply(x, h, S, False, False, *params)
            torch.autograd.backward([z, logdet], [gz, gl])
        tm = ops.collect_kernel_timing()
        ops.enable_kernel_timing(False)
        med = {k: sorted(v)[len(v) // 2] * 1000 for k, v in tm.items()}
        out = [z.detach().clone(), jac.detach().clone(), logdet.detach().clone(), x.grad.clone(), h.grad.clone()] + [p.grad.clone() for p in params]
        print(f"cluster {cs}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in med.items()) + "   (C-ABI calls bracketed by CUDA events, medians of 10)")
        if cs in res:
            continue
        res[cs] = out
    names = ["z", "jac", "logdet", "dx", "dh"] + [f"d{'Wb'[i % 2]}{i // 2}" for i in range(len(params))]
    for n, a, b in zip(names, res[2], res[1]):
        print(f"   {n:7s} cluster 2 vs 1: rel {rel(a, b):.2e}  finite {bool(torch.isfinite(a).all())}")


if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
