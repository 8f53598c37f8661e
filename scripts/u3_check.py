"""Fused strict UMNN forward (gnf_umnn_fwd_tc3) against the layer-wise engine (gnf_umnn_fwd_lw, passes 3 and 0) and a float64
torch evaluation of the same integral: outputs, saved planes / masks, and CUDA-event timings.
usage: u3_check.py [B d S I E]"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnf_b200 as G  # noqa: E402
from gnf_b200 import ops  # noqa: E402
from gnf_b200._lib import lib, ptr, stream_ptr  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def ref64(x, h, Ws, bs, S):
    """float64 torch evaluation of z, jac (CC quadrature over S+1 nodes)."""
    ccw, ccn = ops.cc_weights(S, x.device)
    ccw, ccn = ccw.double(), ccn.double()
    R = x.numel()
    xv = x.double().reshape(R, 1)
    hd = h.double().reshape(R, -1)
    t = xv * (ccn.reshape(1, -1) + 1) / 2                      # [R, S+1]
    inp = torch.cat([t.unsqueeze(-1), hd.unsqueeze(1).expand(-1, S + 1, -1)], -1)
    a = inp
    for l, (W, b) in enumerate(zip(Ws, bs)):
        a = a @ W.double().t() + b.double()
        if l + 1 < len(Ws):
            a = torch.relu(a)
    f = torch.nn.functional.elu(a.squeeze(-1)) + 1.05          # [R, S+1]
    z = (f * ccw.reshape(1, -1)).sum(1) * xv.squeeze(1) / 2 + hd[:, 0]
    return z, f[:, 0]


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
    return e0.elapsed_time(e1) / n * 1000.0


def main(B=100, d=63, S=20, I=150, E=30, nh=3):
    dev = "cuda"
    torch.manual_seed(0)
    dims = [1 + E] + [I] * nh + [1]
    Ws = [torch.empty(dims[l + 1], dims[l], device=dev).uniform_(-1, 1) / dims[l] ** 0.5 for l in range(len(dims) - 1)]
    bs = [torch.empty(dims[l + 1], device=dev).uniform_(-1, 1) / dims[l] ** 0.5 for l in range(len(dims) - 1)]
    x = torch.randn(B, d, device=dev)
    h = torch.randn(B, d, E, device=dev) * 0.5
    R = B * d
    net = ops._mlp_struct(Ws, bs)
    ccw, ccn = ops.cc_weights(S, dev)
    z64, j64 = ref64(x, h, Ws, bs, S)
    out = {}
    only = os.environ.get("U3_ONLY")
    if only:                                                  # profiling runs: one variant, a few launches
        train = int(os.environ.get("U3_TRAIN", "1"))
        saved = torch.zeros(lib().gnf_umnn_lw_saved_floats(C.byref(net), R, S, train), device=dev)
        z = torch.empty(B, d, device=dev); jac = torch.empty(B, d, device=dev); logdet = torch.empty(B, device=dev)
        nb = lib().gnf_umnn_tc3_workspace_bytes(C.byref(net), R)
        ws = torch.empty(nb // 4 + 4, device=dev)
        for _ in range(4):
            rc = lib().gnf_umnn_fwd_tc3(ptr(x), ptr(h), C.byref(net), S, ptr(ccw), ptr(ccn), ptr(z), None, ptr(jac), ptr(logdet),
                                        ptr(saved) if train else None, train, int(only == "tc3o1"), R, d, ptr(ws), nb, stream_ptr())
            assert rc == 0, lib().gnf_last_error()
        torch.cuda.synchronize()
        return
    for train in (1, 0):
        nsaved = lib().gnf_umnn_lw_saved_floats(C.byref(net), R, S, train)
        for name, passes, order in (("lw3", 3, 0), ("lw0", 0, 0), ("tc3", 3, 0), ("tc3o1", 3, 1)):
            z = torch.empty(B, d, device=dev); jac = torch.empty(B, d, device=dev); zrev = torch.empty(B, d, device=dev)
            logdet = torch.empty(B, device=dev)
            saved = torch.zeros(nsaved, device=dev)
            if name.startswith("tc3"):
                nb = lib().gnf_umnn_tc3_workspace_bytes(C.byref(net), R)
                assert nb, lib().gnf_last_error()
                ws = torch.empty(nb // 4 + 4, device=dev)
                def run(sv=saved):
                    rc = lib().gnf_umnn_fwd_tc3(ptr(x), ptr(h), C.byref(net), S, ptr(ccw), ptr(ccn), ptr(z), ptr(zrev), ptr(jac), ptr(logdet),
                                                ptr(sv) if train else None, train, order, R, d, ptr(ws), nb, stream_ptr())
                    assert rc == 0, lib().gnf_last_error()
            else:
                nb = lib().gnf_umnn_lw_workspace_bytes(C.byref(net), R, S, 0)
                ws = torch.empty(nb // 4 + 4, device=dev)
                def run(sv=saved):
                    rc = lib().gnf_umnn_fwd_lw(ptr(x), ptr(h), C.byref(net), S, ptr(ccw), ptr(ccn), ptr(z), ptr(zrev), ptr(jac), ptr(logdet),
                                               ptr(sv), train, passes, R, d, ptr(ws), nb, stream_ptr())
                    assert rc == 0, lib().gnf_last_error()
            run()
            torch.cuda.synchronize()
            us = timed(run)
            out[(name, train)] = (z.clone(), jac.clone(), logdet.clone(), zrev.clone(), saved.clone())
            print(f"train={train} {name:6s}: {us:8.1f} us   z err vs f64 {rel(z.reshape(-1), z64):.2e} (max abs {float((z.reshape(-1).double() - z64).abs().max()):.2e})"
                  f"  jac err {rel(jac.reshape(-1), j64):.2e}  zrev ok {bool(torch.allclose(zrev, z.flip(1)))}")
        # saved buffers: tc3 vs lw3
        a, b = out[("tc3", train)][4], out[("lw3", train)][4]
        nodes = S + 1 + train
        Q = R * nodes
        NP = (I + 31) // 32 * 32
        L = nh
        for l in range(L):
            pa, pb = a[l * Q * NP:(l + 1) * Q * NP], b[l * Q * NP:(l + 1) * Q * NP]
            print(f"   train={train} plane a{l + 1}: rel err tc3 vs lw3 {rel(pa, pb):.2e}  max abs {float((pa - pb).abs().max()):.2e}" + ("" if train else " (not saved in eval)"))
        if train:
            ya, yb = a[L * Q * NP:L * Q * NP + Q], b[L * Q * NP:L * Q * NP + Q]
            print(f"   ysave rel err {rel(ya, yb):.2e}")
            ba = a[L * Q * NP + Q:].view(torch.int32)
            bb = b[L * Q * NP + Q:].view(torch.int32)
            diff = (ba ^ bb)
            nbits = sum(int(((diff >> k) & 1).sum()) for k in range(32))
            print(f"   mask bits differing: {nbits} of {ba.numel() * 32}")
    for k in ("tc3", "tc3o1"):
        print(f"{k} vs lw3 (train): z {rel(out[(k, 1)][0], out[('lw3', 1)][0]):.2e} jac {rel(out[(k, 1)][1], out[('lw3', 1)][1]):.2e} logdet {rel(out[(k, 1)][2], out[('lw3', 1)][2]):.2e}")


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
