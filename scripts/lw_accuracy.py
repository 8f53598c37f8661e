"""Which UMNN engine is closer to the truth?  Per-tensor gradient error of the fused FFMA kernels and of the layer-wise
engine (FFMA / 3xTF32 / TF32 GEMMs) against the CPU oracle evaluated in float64 (and the fp32 oracle for scale)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gnf_b200 as G  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
import gnf_oracle as O  # noqa: E402
import parity  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main(cfg="cfg4", B=100, dev="cuda"):
    spec = G.CONFIGS[cfg]
    model = G.build_from_spec(spec, dev, 0)
    mode = dict(stoch_gate=False)
    parity.set_modes(model, mode)
    x = torch.randn(B, spec["d"], generator=torch.Generator().manual_seed(1))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    modes = [mode] * spec["nb_flow"]
    ospec = {k: v for k, v in spec.items() if k != "A_prior"}
    _, _, _, g32 = O.train_step_grads(x, sd, ospec, modes, None, None)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    _, _, _, g64 = O.train_step_grads(x.double(), sd64, ospec, modes, None, None)
    runs = {"oracle fp32": g32}
    if dev == "cuda":
        for name, engine, gemm, fold in (("fused ffma", "fused", "ffma", 4), ("fused+cond x3", "fused", "tf32x3", 4),
                                         ("lw ffma", "layerwise", "ffma", 4), ("lw x3 f=2", "layerwise", "tf32x3", 2),
                                         ("lw x3 f=4", "layerwise", "tf32x3", 4), ("lw x3 f=inf", "layerwise", "tf32x3", 1 << 20)):
            G.ops.UMNN_ENGINE = engine
            G.ops.set_gemm_mode(gemm)
            G._lib.lib().gnf_tc_gemm_set_fold(fold)
            model.zero_grad()
            z, jac = model(x.to(dev))
            model.loss(z, jac).backward()
            runs[name] = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    if dev == "cuda":
        G._lib.lib().gnf_tc_gemm_set_fold(4)
    keys = [k for k in g64 if g64[k] is not None]
    print(f"{cfg} B={B}: relative L2 error of each gradient tensor vs the float64 oracle")
    print(f"{'tensor':62s}" + "".join(f"{n:>18s}" for n in runs))
    for k in keys:
        print(f"{k:62s}" + "".join(f"{rel(r[k], g64[k]):18.2e}" if r.get(k) is not None else f"{'-':>18s}" for r in runs.values()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cfg4", int(sys.argv[2]) if len(sys.argv) > 2 else 100,
         sys.argv[3] if len(sys.argv) > 3 else "cuda")
