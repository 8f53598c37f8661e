import os
"""Phase timeline of the two-tile tcgen05 UMNN forward kernel (CTA 0), run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import gnf_b200 as G
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import devlib  # noqa: E402  (measurement knobs live in the -DGNF_DEVTOOLS build only)
devlib.install()
lib = G._lib.lib()
model = G.build_from_spec(G.CONFIGS["cfg4"], "cuda", seed=0)
for n in model.getNormalizers():
    n.precision = "tf32"; n.nb_steps = 40
for c in model.getConditioners():
    c.stoch_gate = False
x = torch.randn(2048, 63, device="cuda")
with torch.no_grad():
    model.compute_ll(x)
    buf = torch.zeros(48 * 8, dtype=torch.int64, device="cuda")
    lib.gnf_tc_set_trace(C.c_void_p(buf.data_ptr()))
    model.compute_ll(x)
    torch.cuda.synchronize()
    lib.gnf_tc_set_trace(None)
b = buf.cpu().view(48, 8)
t0 = int(b[0, 0])
names = ["issuer: input staged", "issuer: prev MMA done", "issuer: issued+commit", "epi: own MMA done", "epi: region handed back",
         "epi: reduction done"]
for row in range(16):
    i, g = row // 2, row % 2
    st = [int(v) - t0 if int(v) else -1 for v in b[row, :6]]
    print(f"i={i} wg={g} layer={i%3}: " + "  ".join(f"{n}={v}" for n, v in zip(names, st)))
