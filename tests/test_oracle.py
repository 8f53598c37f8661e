"""CPU tests: the oracle against the reference-generated golden vectors + closed-form KATs."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from helpers import golden_names, load_golden, noises_per_step, rel_err, rel_l2, GOLDEN

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gnf_oracle as O  # noqa: E402


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    c = load_golden(name)
    spec = c["spec"]
    modes = None if c["mode"] is None else [c["mode"]] * spec["nb_flow"]
    loss, z, logdet, grads = O.train_step_grads(c["x"], c["sd"], spec, modes, noises_per_step(c))
    ll = O.normal_log_density(z) + logdet
    assert rel_err(z, c["z"]) < 1e-4 or float((z - c["z"]).abs().max()) < 2e-6
    assert rel_err(ll, c["ll"]) < 1e-5
    assert rel_err(logdet, c["logdet"]) < 1e-4 or float((logdet - c["logdet"]).abs().max()) < 2e-6
    assert abs(float(loss) - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    for k, g in c["grads"].items():
        assert grads[k] is not None, k
        assert rel_l2(grads[k], g) < 2e-5, (k, rel_l2(grads[k], g))


def test_power_trace_golden():
    f = np.load(os.path.join(GOLDEN, "power_trace.npz"))
    keys = sorted({k.rsplit(".", 1)[0] for k in f.files})
    assert len(keys) >= 5
    for key in keys:
        A = torch.from_numpy(f[key + ".A"]).requires_grad_(True)
        d, p, alpha = f[key + ".meta"]
        t = O.power_trace(A, alpha, int(p))
        t.backward()
        assert abs(float(t.detach()) - float(f[key + ".t"])) <= 1e-6 * abs(float(f[key + ".t"])) + 1e-6
        assert rel_l2(A.grad, torch.from_numpy(f[key + ".dA"])) < 1e-6


# ---------------- closed-form known-answer tests ----------------
def test_cc_weights_sum_and_nodes():
    for S in (4, 7, 20, 29, 40, 150):
        w, t = O.cc_weights_nodes(S)
        assert abs(float(w.double().sum()) - 2.) < 1e-5
        assert float(t[0]) == 1.
        assert np.allclose(t.numpy(), np.cos(np.arange(S + 1) * math.pi / S), atol=1e-7)


def test_cc_integrates_polynomials_exactly():
    S = 8
    w, t = O.cc_weights_nodes(S)
    x = 1.7
    nodes = x * (t.double() + 1) / 2
    for deg in range(S + 1):
        est = float((w.double() * nodes ** deg).sum() * x / 2)
        assert abs(est - x ** (deg + 1) / (deg + 1)) < 1e-5 * max(1., x ** (deg + 1))


def test_affine_kat():
    x = torch.randn(3, 4)
    h = torch.zeros(3, 4, 2)
    z, jac = O.affine_normalizer(x, h)
    assert torch.equal(z, x) and torch.equal(jac, torch.ones_like(x))
    h[..., 0] = 9.
    h[..., 1] = -9.
    z, jac = O.affine_normalizer(x, h)
    assert torch.allclose(jac, torch.full_like(x, math.exp(-5.))) and torch.allclose(z, x * math.exp(-5.) + 5.)
    h[..., 1] = 9.
    _, jac = O.affine_normalizer(x, h)
    assert torch.allclose(jac, torch.full_like(x, math.exp(2.)))


def test_normal_log_density_kat():
    assert abs(float(O.normal_log_density(torch.zeros(1, 7))) + 3.5 * math.log(2 * math.pi)) < 1e-5


def test_power_trace_kats():
    assert float(O.power_trace(torch.zeros(5, 5), .2, 5)) == 0.
    A = torch.tril(torch.randn(6, 6), -1)
    assert abs(float(O.power_trace(A, 1 / 6, 6))) < 1e-5
    a, al, p = .8, .5, 5
    A = torch.tensor([[0., a], [a, 0.]])
    want = (1 + al * a * a) ** p + (1 - al * a * a) ** p - 2
    assert abs(float(O.power_trace(A, al, p)) - want) < 1e-5


def test_made_is_autoregressive():
    spec = dict(nb_flow=1, d=5, cond="Autoregressive", hidden=[20, 20], out=3, norm="affine")
    sd = O.init_state_dict(spec, seed=3)
    x = torch.randn(2, 5, requires_grad=True)
    h = O.made_conditioner(x, sd, "steps.0.conditioner", spec)
    for i in range(5):
        g, = torch.autograd.grad(h[:, i, :].sum(), x, retain_graph=True)
        assert float(g[:, i:].abs().max()) == 0.


def test_gate_sigmoid_form():
    p = torch.rand(4, 3, 3)
    u1, u2 = torch.rand(4, 3, 3), torch.rand(4, 3, 3)
    T = .5
    ref = O.gumbel_gate(p, u1, u2, T)
    g1, g2 = -torch.log(-torch.log(u1)), -torch.log(-torch.log(u2))
    alt = torch.sigmoid((torch.log(p + 1e-6) - torch.log(1 - p + 1e-6) + g1 - g2) / T)
    assert float((ref - alt).abs().max()) < 1e-6


def test_umnn_constant_integrand_and_independent_quadrature():
    # last-layer weights zero, bias c  =>  f = elu(c)+1.05 constant  =>  z = f*x + h0
    spec = dict(nb_flow=1, d=3, cond="Coupling", hidden=[4], out=3, norm="monotonic", int_net=[6, 6], nb_steps=6)
    sd = O.init_state_dict(spec, seed=1)
    p = "steps.0.normalizer.integrand_net.net"
    x, h = torch.randn(4, 3), torch.randn(4, 3, 3)
    sdc = dict(sd)
    sdc[f"{p}.4.weight"] = torch.zeros_like(sd[f"{p}.4.weight"])
    sdc[f"{p}.4.bias"] = torch.tensor([.3])
    z, jac = O.monotonic_normalizer(x, h, sdc, p, 3, 6)
    assert torch.allclose(z, 1.35 * x + h[:, :, 0], atol=1e-5)
    assert torch.allclose(jac, torch.full_like(x, 1.35), atol=1e-6)
    # independent evaluation: plain autograd through the quadrature sum (dh must agree; dx differs: Leibniz)
    hq = h.clone().requires_grad_(True)
    z, _ = O.monotonic_normalizer(x, hq, sd, p, 3, 6)
    gz = torch.randn_like(z)
    gh, = torch.autograd.grad(z, hq, gz)
    w, t = O.cc_weights_nodes(6)
    hq2 = h.clone().requires_grad_(True)
    xr, hr = x.reshape(-1), hq2.reshape(-1, 3)
    acc = sum(w[k] * O.integrand(xr * (t[k] + 1) / 2, hr, sd, p, 3) for k in range(7))
    z2 = (acc * xr / 2).view(4, 3) + hq2[:, :, 0]
    gh2, = torch.autograd.grad(z2, hq2, gz)
    assert torch.allclose(z, z2, atol=1e-5)
    assert rel_l2(gh, gh2) < 1e-5
