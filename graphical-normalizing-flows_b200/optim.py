"""Optimizer of the reference's training loops on the sm_100a library: torch.optim.Adam(params, lr, weight_decay)
(UCIExperiments.py:100, ToyExperiments.py:59) as ONE multi-tensor kernel launch per step (csrc/optim.cu)."""
import ctypes as C

import torch

from . import _lib as L
from . import ops


class FusedAdam(torch.optim.Optimizer):
    """Same update rule and hyper-parameter names as torch.optim.Adam (L2 weight decay, no amsgrad, no maximize).  The step
    count lives on the device, so `step()` can be captured in a CUDA graph (GraphedTrainStep)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._table_key, self._table = None, None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                L.require(p, "parameter"), L.require(p.grad, "gradient")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            if not entries:
                continue
            if "step" not in group:
                group["step"] = torch.zeros(1, dtype=torch.int64, device=entries[0][0].device)
            key = (gi,) + tuple((p.data_ptr(), g.data_ptr(), p.numel()) for p, g, _, _ in entries)
            if key != self._table_key:
                arr = (L.AdamTensorT * len(entries))()
                for i, (p, g, m, v) in enumerate(entries):
                    arr[i].param, arr[i].grad, arr[i].exp_avg, arr[i].exp_avg_sq, arr[i].numel = \
                        p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()
                self._table_key, self._table = key, arr
            b1, b2 = group["betas"]
            ops._call("gnf_adam_step", self._table, len(entries), L.ptr(group["step"]), float(group["lr"]), float(b1), float(b2),
                      float(group["eps"]), float(group["weight_decay"]), L.stream_ptr())
            ops._count()
            ops.counter_add(group["step"], 1)
        return loss
