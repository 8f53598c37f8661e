import os
"""TMEM-read bandwidth / TF32 MMA rate / overlap probe (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
lib = G._lib.lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
iters = 200
for mode, name in ((1, "tcgen05.ld only"), (2, "mma only"), (3, "both concurrently")):
    out.zero_()
    G._lib.check(lib.gnf_tc_probe(mode, iters, C.c_void_p(out.data_ptr()), G._lib.stream_ptr()))
    torch.cuda.synchronize()
    ld, mma = int(out[0]), int(out[1])
    msg = f"{name:20s}"
    if mode & 1:
        by = iters * 128 * 256 * 4
        msg += f" ld: {ld} clk, {by / ld:.1f} B/clk (128 lanes x 256 cols x 4 B x {iters})"
    if mode & 2:
        msg += f" mma: {mma} clk, {mma / (iters * 19):.1f} clk per M128xN160xK8 TF32 MMA"
    print(msg)
