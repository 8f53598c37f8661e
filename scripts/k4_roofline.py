"""HBM roofline of the K4 kernels (AffineNormalizer + log-det, base density: gnf_affine_{fwd,bwd}, gnf_normal_ll_{fwd,bwd}) at sizes
beyond the 126 MB L2: algorithmic bytes (forward 16*B*d + 8*B) / CUDA-event duration against the measured copy
bandwidth of MEASURED_PEAKS.json.  Prints one JSON object (committed under profiles/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnf_b200 as G  # noqa: E402


def timed(fn, n=20):
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


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.)
    out = {"peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)", "cases": []}
    for name, B, d, H in (("cfg5 shape x 256 batches (d=784, H=2)", 25600, 784, 2), ("cfg1 shape (d=2, H=150: h rows are 600 B apart)", 4000000, 2, 150),
                          ("d=63, H=2", 400000, 63, 2)):
        x = torch.randn(B, d, device="cuda")
        h = torch.randn(B, d, H, device="cuda")
        norm = G.AffineNormalizer()
        dens = G.NormalLogDensity().to("cuda")
        xg, hg = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
        G.ops.enable_kernel_timing(True)
        for _ in range(12):
            z, jac = norm(xg, hg)
            (dens(z).sum() + z.sum() + jac.sum()).backward()
            xg.grad = None
        t = G.ops.collect_kernel_timing()
        G.ops.enable_kernel_timing(False)
        case = {"case": name, "B": B, "d": d, "H": H, "kernels": []}
        # algorithmic bytes: fwd reads x, h[...,0:2], writes z, jac(+mask); bwd reads x, h01, mask, gz, writes gx, gh[...,0:H]
        by = {"gnf_affine_fwd": 16. * B * d + 8. * B + 4. * B * d, "gnf_affine_bwd": (16. + 4. * H) * B * d + 8. * B,
              "gnf_normal_ll_fwd": 4. * B * d + 8. * B, "gnf_normal_ll_bwd": 8. * B * d + 8. * B}
        for k, nb in by.items():
            if k in t:
                ms = sorted(t[k])[len(t[k]) // 2]
                case["kernels"].append({"bound": "hbm", "kernel": k, "bytes": nb, "ms": ms, "achieved": nb / ms / 1e6, "peak": peak, "unit": "GB/s",
                                        "frac": nb / ms / 1e6 / peak})
        out["cases"].append(case)
        del x, h, xg, hg
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
