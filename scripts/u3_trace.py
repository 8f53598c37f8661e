"""Warp-role timeline of the fused strict UMNN forward (dev build): SM-clock stamps of CTA 0."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402

lib = devlib.install()
import gnf_b200 as G  # noqa: E402
from gnf_b200 import ops  # noqa: E402
from gnf_b200._lib import ptr, stream_ptr  # noqa: E402


def main(B=100, d=63, S=20, I=150, E=30, nh=3, train=1, order=0, debug=0):
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
    z = torch.empty(B, d, device=dev); jac = torch.empty(B, d, device=dev); logdet = torch.empty(B, device=dev)
    nsaved = lib.gnf_umnn_lw_saved_floats(C.byref(net), R, S, train)
    saved = torch.zeros(nsaved, device=dev)
    nb = lib.gnf_umnn_tc3_workspace_bytes(C.byref(net), R)
    ws = torch.empty(nb // 4 + 4, device=dev)
    trace = torch.zeros(4 * 256, dtype=torch.int64, device=dev)

    def run():
        rc = lib.gnf_umnn_fwd_tc3(ptr(x), ptr(h), C.byref(net), S, ptr(ccw), ptr(ccn), ptr(z), None, ptr(jac), ptr(logdet),
                                  ptr(saved) if train else None, train, order, R, d, ptr(ws), nb, stream_ptr())
        assert rc == 0, lib.gnf_last_error()
    lib.gnf_umnn_tc3_set_debug(debug)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"debug={debug}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per call")
    lib.gnf_umnn_tc3_set_trace(ptr(trace))
    run()
    torch.cuda.synchronize()
    lib.gnf_umnn_tc3_set_trace(None)
    t = trace.cpu().view(4, 256)
    t0 = int(t[t > 0].min())
    names = ["issuer", "epi warp0", "producer", "epi warp4"]
    print(f"B={B} d={d} S={S} I={I} train={train} order={order}: SM clocks relative to the first stamp")
    for i in range(4):
        row = [int(v) - t0 for v in t[i] if v > 0]
        print(f"{names[i]:10s} n={len(row):3d} {row[:64]}")


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
