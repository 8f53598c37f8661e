"""Flow composition and factories with the reference's API
(models/NormalizingFlow.py, models/NormalizingFlowFactories.py)."""
from math import pi

import torch
import torch.nn as nn

from . import ops
from .conditioners import Conditioner, DAGConditioner
from .normalizers import Normalizer


class NormalizingFlow(nn.Module):
    """Abstract API (NormalizingFlow.py:7-58)."""

    def __init__(self):
        super().__init__()

    def forward(self, x, context=None):
        pass

    def constraintsLoss(self):
        pass

    def DAGness(self):
        pass

    def step(self, epoch_number, loss_avg):
        pass

    def getConditioners(self):
        pass

    def isInvertible(self):
        pass

    def getNormalizers(self):
        pass

    def invert(self, z, context=None):
        pass


class NormalizingFlowStep(NormalizingFlow):
    """NormalizingFlow.py:61-107."""

    def __init__(self, conditioner: Conditioner, normalizer: Normalizer):
        super().__init__()
        self.conditioner = conditioner
        self.normalizer = normalizer

    def forward_fused(self, x, context=None, want_rev=False):
        h = self.conditioner(x, context)
        out = self.normalizer.forward_fused(x, h, want_rev)
        z, _, logdet, zrev = out
        return z, logdet, zrev

    def forward(self, x, context=None):
        z, logdet, _ = self.forward_fused(x, context)
        return z, logdet

    def constraintsLoss(self):
        if type(self.conditioner) is DAGConditioner:
            return self.conditioner.loss()
        return 0.

    def DAGness(self):
        if type(self.conditioner) is DAGConditioner:
            return [self.conditioner.get_power_trace()]
        return [0.]

    def step(self, epoch_number, loss_avg):
        if type(self.conditioner) is DAGConditioner:
            self.conditioner.step(epoch_number, loss_avg)

    def getConditioners(self):
        return [self.conditioner]

    def getNormalizers(self):
        return [self.normalizer]

    def isInvertible(self):
        for conditioner in self.getConditioners():
            if not conditioner.is_invertible:
                return False
        return True

    def invert(self, z, context=None):
        # fixed-point iteration over the DAG depth (NormalizingFlow.py:98-107), without the progress prints
        x = torch.zeros_like(z)
        with torch.no_grad():
            for _ in range(self.conditioner.depth() + 1):
                h = self.conditioner(x, context)
                x_prev = x
                x = self.normalizer.inverse_transform(z, h, context)
                if torch.norm(x - x_prev) == 0.:
                    break
        return x


class FCNormalizingFlow(NormalizingFlow):
    """NormalizingFlow.py:110-169 (+ compute_ll, which the reference's drivers call but never define: quirk Q10)."""

    def __init__(self, steps, z_log_density):
        super().__init__()
        self.steps = nn.ModuleList()
        self.z_log_density = z_log_density
        for step in steps:
            self.steps.append(step)

    def forward(self, x, context=None):
        if x.dim() == 2 and x.shape[0] == 0:
            # empty batch: nothing to launch (the reference returns empty tensors too; keep the graph connected to x)
            return x * 1., x.sum(1)
        jac_tot = None                       # the reference starts from 0. (NormalizingFlow.py:119): no `0. + tensor` launch for the first step
        n = len(self.steps)
        z = None
        for k, step in enumerate(self.steps):
            # the column reversal feeding the next step is an epilogue of this step's normalizer kernel
            z, jac, zrev = step.forward_fused(x, context, want_rev=(k < n - 1))
            x = zrev
            jac_tot = jac if jac_tot is None else jac_tot + jac
        return z, jac_tot

    def compute_ll(self, x, context=None):
        """ll [B], z [B,d] — the closure of ToyExperiments.py:134-137 / UCIExperiments.py:159-160."""
        z, jac = self.forward(x, context)
        if x.dim() == 2 and x.shape[0] == 0:
            return jac, z
        if isinstance(self.z_log_density, NormalLogDensity):
            return ops.NormalLLFn.apply(z.contiguous(), jac.contiguous()), z
        return self.z_log_density(z) + jac, z

    def constraintsLoss(self):
        loss = None
        for step in self.steps:
            l = step.constraintsLoss()
            if torch.is_tensor(l):
                loss = l if loss is None else loss + l      # no `0. + tensor` launch for the first term
            elif l != 0.:
                loss = l if loss is None else loss + l
        return 0. if loss is None else loss

    def DAGness(self):
        dagness = []
        for step in self.steps:
            dagness += step.DAGness()
        return dagness

    def step(self, epoch_number, loss_avg):
        for step in self.steps:
            step.step(epoch_number, loss_avg)

    def loss(self, z, jac, constraint=None):
        """NormalizingFlow.py:144-146.  constraint: a constraintsLoss() value the caller already has (GraphedTrainStep computes it on a
        side branch of the captured step, next to the forward); None = compute it here, as the reference does."""
        c = self.constraintsLoss() if constraint is None else constraint
        if isinstance(self.z_log_density, NormalLogDensity) and z.dim() == 2 and z.shape[0] > 0:
            # constraint - mean(jac + z_log_density(z)) as one kernel per direction
            tensor_c = torch.is_tensor(c) and c.dim() == 0 and c.dtype == z.dtype and c.device == z.device
            out = ops.NllLossFn.apply(z.contiguous(), jac.contiguous(), c if tensor_c else None)
            return out if tensor_c else c + out
        log_p_x = jac + self.z_log_density(z)
        return c - log_p_x.mean()

    def getNormalizers(self):
        normalizers = []
        for step in self.steps:
            normalizers += step.getNormalizers()
        return normalizers

    def getConditioners(self):
        conditioners = []
        for step in self.steps:
            conditioners += step.getConditioners()
        return conditioners

    def isInvertible(self):
        for conditioner in self.getConditioners():
            if not conditioner.is_invertible:
                return False
        return True

    def invert(self, z, context=None):
        """Inverse of forward().  The reference's own multi-step invert is wrong (quirk Q9: it indexes
        steps[-0] first and never undoes the column reversal); this one visits the steps last-to-first and
        undoes the reversal, and is identical to the reference for 1-step flows."""
        n = len(self.steps)
        for k in range(n - 1, -1, -1):
            x = self.steps[k].invert(z, context)
            if k > 0:
                z = x.flip(1)
        return x


class NormalLogDensity(nn.Module):
    """NormalizingFlowFactories.py:10-16."""

    def __init__(self):
        super().__init__()
        self.register_buffer("pi", torch.tensor(pi))

    def forward(self, z):
        return ops.NormalLLFn.apply(z.contiguous(), None)


def buildFCNormalizingFlow(nb_steps, conditioner_type, conditioner_args, normalizer_type, normalizer_args):
    """NormalizingFlowFactories.py:19-32."""
    flow_steps = []
    for step in range(nb_steps):
        conditioner = conditioner_type(**conditioner_args)
        normalizer = normalizer_type(**normalizer_args)
        flow_steps.append(NormalizingFlowStep(conditioner, normalizer))
    return FCNormalizingFlow(flow_steps, NormalLogDensity())


def MNIST_A_prior(in_size, kernel):
    """Local-window adjacency prior (NormalizingFlowFactories.py:35-46): pixel p depends on the pixels of the
    (2k+1)x(2k+1) window around it, zero diagonal."""
    r = torch.arange(in_size)
    row = r.view(-1, 1).expand(in_size, in_size).reshape(-1)      # row index of each flattened pixel
    col = r.view(1, -1).expand(in_size, in_size).reshape(-1)
    dr = (row.view(-1, 1) - row.view(1, -1)).abs()
    dc = (col.view(-1, 1) - col.view(1, -1)).abs()
    A = ((dr <= kernel) & (dc <= kernel)).float()
    A.fill_diagonal_(0.)
    return A
