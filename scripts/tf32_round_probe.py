"""Does tcgen05 kind::tf32 truncate or round the fp32 operand bits it reads? (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnf_b200 as G
for mode in (0, 1):
    A = torch.zeros(128, 8, device="cuda"); W = torch.zeros(16, 8, device="cuda")
    vals = [1 + 2**-11 + 2**-12, 1 + 2**-11, 1 + 2**-11 - 2**-20, 1 + 2**-10 + 2**-11, -(1 + 2**-11 + 2**-12), 1 + 2**-12]
    for i, v in enumerate(vals):
        A[i, 0] = v
    W[0, 0] = 1.0
    W[1, 0] = 1 + 2**-11 + 2**-12   # B-side rounding, A = exact ones
    A[64:, 0] = 1.0
    C = G.ops.tc_selftest(A, W, mode)
    print("mode", mode, "A-side:", [(f"{v:.10f}", f"{float(C[i,0]):.10f}") for i, v in enumerate(vals)])
    print("        B-side:", f"{float(C[64,1]):.10f}", "(trunc -> 1.0000000000, RN -> 1.0009765625)")
