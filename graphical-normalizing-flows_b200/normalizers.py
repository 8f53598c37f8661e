"""Normalizer plugins with the reference's names / signatures / state_dict keys
(models/Normalizers/*.py), rebinding ``forward`` to the sm_100a kernels."""
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class Normalizer(nn.Module):
    """Plugin ABC (models/Normalizers/Normalizer.py:4-28): forward(x, h, context) -> (z [B,d], jac [B,d])."""

    def __init__(self):
        super().__init__()

    def forward(self, x, h, context=None):
        pass

    def inverse_transform(self, z, h, context=None):
        pass


class AffineNormalizer(Normalizer):
    """models/Normalizers/AffineNormalizer.py:5-17."""

    def __init__(self):
        super().__init__()

    def forward_fused(self, x, h, want_rev=False):
        """(z, jac, logdet, zrev) in one kernel: the step's log(jac).sum(1) and the next step's column
        reversal are epilogues of the same pass (NormalizingFlow.py:70,120)."""
        z, jac, logdet, zrev = ops.AffineFn.apply(x.contiguous(), h, want_rev)
        return z, jac, logdet, (zrev if want_rev else None)

    def forward(self, x, h, context=None):
        z, jac, _, _ = self.forward_fused(x, h)
        return z, jac

    def inverse_transform(self, z, h, context=None):
        # closed form (AffineNormalizer.py:14-17); sampling path, SURVEY.md §8f rank 2
        mu, sigma = h[:, :, 0].clamp_(-5., 5.), torch.exp(h[:, :, 1].clamp_(-5., 2.))
        return (z - mu) / sigma


class ELUPlus(nn.Module):
    """ELU(x) + 1.05 (MonotonicNormalizer.py:12-18); parameter-free, evaluated inside the UMNN kernel."""

    def __init__(self):
        super().__init__()
        self.elu = nn.ELU()

    def forward(self, x):
        return self.elu(x) + 1.05


class IntegrandNet(nn.Module):
    """Parameter container of the positive integrand MLP (MonotonicNormalizer.py:21-38)."""

    def __init__(self, hidden, cond_in):
        super().__init__()
        l1 = [1 + cond_in] + list(hidden)
        l2 = list(hidden) + [1]
        layers = []
        for h1, h2 in zip(l1, l2):
            layers += [nn.Linear(h1, h2), nn.ReLU()]
        layers.pop()
        layers.append(ELUPlus())
        self.net = nn.Sequential(*layers)

    def linear_params(self):
        params = []
        for m in self.net:
            if isinstance(m, nn.Linear):
                params += [m.weight, m.bias]
        return params

    def forward(self, x, h):
        """f(x; h) with the reference's argument layout (x [N,d], h [N, E*d] with h[n, k*d+i]) — evaluated by the
        UMNN kernel's Jacobian output (node 0 of the quadrature is the upper limit x itself)."""
        N, d = x.shape
        h3 = h.view(N, -1, d).permute(0, 2, 1).contiguous()
        _, jac, _, _ = ops.UmnnFn.apply(x.contiguous(), h3, 1, False, False, *self.linear_params())
        return jac


class MonotonicNormalizer(Normalizer):
    """models/Normalizers/MonotonicNormalizer.py:41-83."""

    def __init__(self, integrand_net, cond_size, nb_steps=20, solver="CC"):
        super().__init__()
        if type(integrand_net) is list:
            self.integrand_net = IntegrandNet(integrand_net, cond_size)
        else:
            raise NotImplementedError("MonotonicNormalizer(integrand_net=<nn.Module>): only the list form (IntegrandNet) is "
                                      "covered by the fused UMNN kernel; there is no eager fallback")
        self.solver = solver
        self.nb_steps = nb_steps
        # "strict": fp32 FFMA kernels (ll 1e-4, gradients 1e-3).  "tf32": forward on the tensor cores with
        # single-pass TF32 operands (ll 2e-3); gradients still come from the strict backward kernel.
        self.precision = "strict"

    def forward_fused(self, x, h, want_rev=False):
        if self.solver not in ("CC", "CCParallel"):
            return None
        # "CC" (sequential) and "CCParallel" (batched) are the same quadrature; one kernel serves both.
        if self.precision not in ("strict", "tf32"):
            raise ValueError(f"MonotonicNormalizer.precision must be 'strict' or 'tf32', got {self.precision!r}")
        z, jac, logdet, zrev = ops.UmnnFn.apply(x.contiguous(), h.contiguous(), int(self.nb_steps), want_rev,
                                                self.precision == "tf32", *self.integrand_net.linear_params())
        return z, jac, logdet, (zrev if want_rev else None)

    def forward(self, x, h, context=None):
        out = self.forward_fused(x, h)
        if out is None:
            return None
        return out[0], out[1]

    # inverse_transform runs the whole bisection in one kernel launch (gnf_umnn_invert, FFMA tiles that hold every quadrature node
    # of their rows: S + 1 <= 64) -- 1.9x the loop of 20 launches of the same FFMA forward (8.1 vs 15.3 ms at cfg4's shape).  Where
    # the forward itself runs on the fused tcgen05 kernel (large batches in a tensor-core GEMM mode) 20 launches of THAT kernel
    # are faster still (4.6 ms, profiles/r02t_invert_bench.txt): 'auto' picks by the same rule as the forward.  True / False force.
    fused_inverse = "auto"

    def inverse_transform(self, z, h, context=None):
        # 20-step bisection on [-20, 20] (MonotonicNormalizer.py:69-83); sampling path, SURVEY.md §8f rank 2
        fused = self.fused_inverse
        if fused == "auto":
            fused = not ops.umnn_forward_on_tensor_cores(self.integrand_net.linear_params(), z.numel(), int(self.nb_steps), z.device)
        if fused and self.solver in ("CC", "CCParallel") and int(self.nb_steps) + 1 <= 64 and (z.is_cuda or L._SIMULATOR):
            with torch.no_grad():
                ws = self.integrand_net.linear_params()
                return ops.umnn_invert(z.contiguous(), h.contiguous(), int(self.nb_steps), ws[0::2], ws[1::2], 20, -20., 20.)
        x_max = torch.ones_like(z) * 20
        x_min = -torch.ones_like(z) * 20
        with torch.no_grad():
            for _ in range(20):
                x_middle = (x_max + x_min) / 2
                z_middle, _ = self.forward(x_middle, h, context)
                left = (z_middle > z).float()
                right = 1 - left
                x_max = left * x_middle + right * x_max
                x_min = right * x_middle + left * x_min
        return (x_max + x_min) / 2
