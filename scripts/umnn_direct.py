"""Direct accuracy probe of the UMNN normalizer engines at the cfg4 integrand shape: random x / h / weights, a random
linear functional of (z, log jac) as the loss, gradients against the oracle's quadrature evaluated in float64 on the GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gnf_b200 as G  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
import gnf_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main(B=100, d=63, hscale=1.0, S=20):
    dev = "cuda"
    torch.manual_seed(0)
    norm = G.MonotonicNormalizer([150, 150, 150], 30, nb_steps=S, solver="CC").to(dev)
    x = torch.randn(B, d, device=dev)
    h = (torch.randn(B, d, 30, device=dev) * hscale)
    cz = torch.randn(B, d, device=dev)
    cj = torch.randn(B, d, device=dev)

    def loss_of(z, jac):
        return (z * cz).sum() + (torch.log(jac) * cj).sum()

    # float64 reference (oracle quadrature, torch on the GPU)
    sd64 = {"n." + k: v.detach().double() for k, v in norm.state_dict().items()}
    x64, h64 = x.double().requires_grad_(True), h.double().requires_grad_(True)
    for v in sd64.values():
        v.requires_grad_(True)
    n_layers = len([k for k in sd64 if k.endswith("weight")])
    z, jac = O.monotonic_normalizer(x64, h64, sd64, "n.integrand_net.net", n_layers, S)
    ((z * cz.double()).sum() + (torch.log(jac) * cj.double()).sum()).backward()
    ref = {"x": x64.grad, "h": h64.grad}
    ref.update({k[2:]: v.grad for k, v in sd64.items() if v.grad is not None})
    print(f"B={B} d={d} hscale={hscale}: relative L2 error vs float64")
    rows = {}
    for name, engine, gemm, fold in (("fused", "fused", "ffma", 4), ("lw ffma", "layerwise", "ffma", 4), ("lw x3 f=1", "layerwise", "tf32x3", 1),
                                     ("lw x3 f=2", "layerwise", "tf32x3", 2), ("lw x3 f=4", "layerwise", "tf32x3", 4),
                                     ("lw x3 f=8", "layerwise", "tf32x3", 8), ("lw x3 f=inf", "layerwise", "tf32x3", 1 << 20)):
        G.ops.UMNN_ENGINE = engine
        G.ops.set_gemm_mode(gemm)
        G._lib.lib().gnf_tc_gemm_set_fold(fold)
        xg, hg = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
        norm.zero_grad()
        z32, j32 = norm(xg, hg)
        loss_of(z32, j32).backward()
        got = {"x": xg.grad, "h": hg.grad}
        got.update({k: p.grad for k, p in norm.named_parameters()})
        got["z(fwd)"], got["jac(fwd)"] = z32.detach(), j32.detach()
        rows[name] = got
    ref["z(fwd)"], ref["jac(fwd)"] = z.detach(), jac.detach()
    G._lib.lib().gnf_tc_gemm_set_fold(4)
    print(f"{'tensor':34s}" + "".join(f"{n:>14s}" for n in rows))
    for k in ref:
        print(f"{k:34s}" + "".join(f"{rel(r[k], ref[k]):14.2e}" for r in rows.values()))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 100, int(a[1]) if len(a) > 1 else 63, float(a[2]) if len(a) > 2 else 1.0)
