"""MonotonicNormalizer.inverse_transform: the fused bisection kernel (gnf_umnn_invert) against the reference's loop of 20 forward
passes (each a launch of the CUDA forward kernel + six elementwise torch ops).  usage: invert_bench.py [B d S I E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnf_b200 as G  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main(B=100, d=63, S=20, I=150, E=30):
    torch.manual_seed(0)
    norm = G.MonotonicNormalizer([I, I, I], E, nb_steps=S, solver="CC").to("cuda")
    h = torch.randn(B, d, E, device="cuda") * .5
    x = torch.randn(B, d, device="cuda") * 2
    for mode in ("ffma", "auto"):
        G.ops.set_gemm_mode(mode)
        with torch.no_grad():
            z, _ = norm(x, h)
            norm.fused_inverse = True
            tf = timed(lambda: norm.inverse_transform(z, h))
            xf = norm.inverse_transform(z, h)
            norm.fused_inverse = False
            tl = timed(lambda: norm.inverse_transform(z, h))
            xl = norm.inverse_transform(z, h)
            norm.fused_inverse = "auto"
            ta = timed(lambda: norm.inverse_transform(z, h))
        print(f"B={B} d={d} S={S} I={I} gemm mode {mode:5s}: fused search {tf:7.3f} ms   loop of 20 forward passes {tl:7.3f} ms   auto {ta:7.3f} ms   "
              f"max |x_fused - x| {float((xf - x).abs().max()):.2e}  max |x_loop - x| {float((xl - x).abs().max()):.2e}")


if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
