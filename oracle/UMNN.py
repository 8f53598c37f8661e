"""CPU restatement of the third-party package ``UMNN==1.0`` (TEST INFRASTRUCTURE ONLY).

The reference pins ``UMNN==1.0`` (``/root/reference/requirements.txt:3``) and imports
``NeuralIntegral`` / ``ParallelNeuralIntegral`` from it at
``models/Normalizers/MonotonicNormalizer.py:2`` (call sites ``:58-63``).  The package is
not vendored under ``/root/reference``, is not installed in this image and cannot be
downloaded (no network), so its published algorithm (Clenshaw-Curtis quadrature of a
positive integrand network with a recompute-based backward, Wehenkel & Louppe 2019,
"Unconstrained Monotonic Neural Networks") is restated here from SURVEY.md Appendix B.

PARITY UNPINNED at this boundary: the reference holds no test, golden vector or fixture
for the integral, so this restatement is checked only against closed-form known answers
(tests/test_oracle.py) and against an independent autograd-through-quadrature evaluation.

Only ``tests/``, ``__graft_entry__.smoke()``, the golden-vector generator and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module.  The
product package (``graphical-normalizing-flows_b200``) never does.
"""
import math

import numpy as np
import torch


def _flatten(sequence):
    flat = [p.contiguous().view(-1) for p in sequence]
    return torch.cat(flat) if len(flat) > 0 else torch.tensor([])


def compute_cc_weights(nb_steps):
    """Clenshaw-Curtis weights / nodes on [-1, 1] in float64, cast to fp32 (App. B)."""
    k = np.arange(0, nb_steps + 1).reshape(-1, 1)
    lam = np.cos(k @ k.T * math.pi / nb_steps)
    lam[:, 0] = .5
    lam[:, -1] = .5 * lam[:, -1]
    lam = lam * 2 / nb_steps
    W = np.arange(0, nb_steps + 1).reshape(-1, 1).astype(np.float64)
    W[np.arange(1, nb_steps + 1, 2)] = 0
    W = 2 / (1 - W ** 2)
    W[0] = 1
    W[np.arange(1, nb_steps + 1, 2)] = 0
    cc_weights = torch.tensor(lam.T @ W).float()
    steps = torch.tensor(np.cos(np.arange(0, nb_steps + 1).reshape(-1, 1) * math.pi / nb_steps)).float()
    return cc_weights, steps


def _integrand_grads(x, integrand, h, cotangent):
    with torch.enable_grad():
        h = h.detach().requires_grad_(True)
        f = integrand.forward(x, h)
        g_param = _flatten(torch.autograd.grad(f, integrand.parameters(), cotangent,
                                               retain_graph=True))
        g_h = _flatten(torch.autograd.grad(f, h, cotangent, retain_graph=True))
    return g_param, g_h


def _integrate_parallel(x0, nb_steps, step_sizes, integrand, h, compute_grad=False, x_tot=None):
    cc_weights, steps = compute_cc_weights(nb_steps)
    cc_weights, steps = cc_weights.to(x0.device), steps.to(x0.device)
    xT = x0 + nb_steps * step_sizes
    B, d = x0.shape
    x0_t = x0.unsqueeze(1).expand(-1, nb_steps + 1, -1)
    xT_t = xT.unsqueeze(1).expand(-1, nb_steps + 1, -1)
    h_steps = h.unsqueeze(1).expand(-1, nb_steps + 1, -1)
    steps_t = steps.unsqueeze(0).expand(B, -1, d)
    X_steps = x0_t + (xT_t - x0_t) * (steps_t + 1) / 2
    X_steps = X_steps.contiguous().view(-1, d)
    h_steps = h_steps.contiguous().view(-1, h.shape[1])
    if not compute_grad:
        dzs = integrand(X_steps, h_steps).view(B, nb_steps + 1, -1)
        dzs = dzs * cc_weights.unsqueeze(0).expand(dzs.shape)
        z_est = dzs.sum(1)
        return z_est * (xT - x0) / 2
    x_tot = x_tot * (xT - x0) / 2
    x_tot_steps = x_tot.unsqueeze(1).expand(-1, nb_steps + 1, -1) * \
        cc_weights.unsqueeze(0).expand(B, -1, d)
    x_tot_steps = x_tot_steps.contiguous().view(-1, d)
    g_param, g_h = _integrand_grads(X_steps, integrand, h_steps, x_tot_steps)
    return g_param, g_h.view(B, nb_steps + 1, -1).sum(1)


def _integrate_sequential(x0, nb_steps, step_sizes, integrand, h, compute_grad=False, x_tot=None):
    cc_weights, steps = compute_cc_weights(nb_steps)
    cc_weights, steps = cc_weights.to(x0.device), steps.to(x0.device)
    xT = x0 + nb_steps * step_sizes
    if not compute_grad:
        z = 0.
        for i in range(nb_steps + 1):
            x = x0 + (xT - x0) * (steps[i] + 1) / 2
            z = z + cc_weights[i] * integrand(x, h)
        return z * (xT - x0) / 2
    g_param, g_h = 0., 0.
    for i in range(nb_steps + 1):
        x = x0 + (xT - x0) * (steps[i] + 1) / 2
        dg_param, dg_h = _integrand_grads(x, integrand, h, cc_weights[i] * x_tot * (xT - x0) / 2)
        g_param = g_param + dg_param
        g_h = g_h + dg_h
    return g_param, g_h


def _make_integral(integrate):
    class _Integral(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x0, x, integrand, flat_params, h, nb_steps=20):
            with torch.no_grad():
                x_tot = integrate(x0, nb_steps, (x - x0) / nb_steps, integrand, h, False)
                ctx.integrand = integrand
                ctx.nb_steps = nb_steps
                ctx.save_for_backward(x0.clone(), x.clone(), h)
            return x_tot

        @staticmethod
        def backward(ctx, grad_output):
            x0, x, h = ctx.saved_tensors
            integrand, nb_steps = ctx.integrand, ctx.nb_steps
            g_param, g_h = integrate(x0, nb_steps, x / nb_steps, integrand, h, True, grad_output)
            x_grad = integrand(x, h)
            x0_grad = integrand(x0, h)
            return -x0_grad * grad_output, x_grad * grad_output, None, g_param, g_h.view(h.shape), None
    return _Integral


NeuralIntegral = _make_integral(_integrate_sequential)
NeuralIntegral.__name__ = "NeuralIntegral"
ParallelNeuralIntegral = _make_integral(_integrate_parallel)
ParallelNeuralIntegral.__name__ = "ParallelNeuralIntegral"


class UMNNMAFFlow(torch.nn.Module):
    """Placeholder: imported (never used) by the reference's image driver."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("UMNNMAFFlow is outside the hot path (SURVEY.md §2 row 6)")
