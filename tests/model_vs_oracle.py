"""Seeded product model vs the CPU oracle on the same inputs / weights / gate noise (any size)."""
import os
import sys

import torch

import gnf_b200 as G
from helpers import rel_err, rel_l2
import parity

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gnf_oracle as O  # noqa: E402

CONFIGS = G.CONFIGS


def build(spec, device, seed=0):
    return G.build_from_spec(spec, device, seed)


def compare(spec, B, device, mode=None, seed=0, nb_steps=None, train=True, scaleA=None):
    """Returns a dict of relative errors (ll per sample: max relative; gradients: per-tensor relative L2)."""
    model = build(spec, device, seed)
    ospec = {k: v for k, v in spec.items() if k != "A_prior"}
    if scaleA is not None:
        with torch.no_grad():
            for c in model.getConditioners():
                c.A.mul_(scaleA)
    parity.set_modes(model, mode)
    if nb_steps is not None:
        for n in model.getNormalizers():
            n.nb_steps = nb_steps
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, spec["d"], generator=g).to(device)
    model.zero_grad()
    z, jac = model(x)
    ll = model.z_log_density(z) + jac
    noises = None
    if spec["cond"] == "DAG":
        noises = []
        for c in model.getConditioners():
            gs = c._last_gate
            noises.append(tuple(n.cpu() for n in G.ops.dag_dump_noise(gs, B, spec["d"], x.device))
                          if gs.mode != G._lib.GATE_TABLE else None)
        if all(n is None for n in noises):
            noises = None
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    modes = None if mode is None else [mode] * spec["nb_flow"]
    rep = {}
    if train:
        loss = model.loss(z, jac)
        loss.backward()
        loss_o, z_o, jac_o, grads_o = O.train_step_grads(x.cpu(), sd, ospec, modes, noises, nb_steps)
        rep["loss"] = abs(float(loss.detach()) - float(loss_o)) / max(abs(float(loss_o)), 1e-6)
        params = dict(model.named_parameters())
        for k, go in grads_o.items():
            if go is None:
                continue
            assert params[k].grad is not None, k
            rep["grad." + k] = rel_l2(params[k].grad.detach().cpu(), go)
    else:
        with torch.no_grad():
            z_o, jac_o = O.flow_forward(x.cpu(), sd, ospec, modes, noises, nb_steps)
    ll_o = O.normal_log_density(z_o) + jac_o
    rep["ll"] = rel_err(ll.detach().cpu(), ll_o)
    rep["z_abs"] = float((z.detach().cpu() - z_o).abs().max())
    return rep
