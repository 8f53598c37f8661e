"""Parity harness shared by the GPU tests (real library, CUDA tensors) and the simulator tests
(host SIMT build of the same kernels, CPU tensors): build the product model from a golden case's spec,
load the reference's state_dict, replay the reference's noise, compare against the reference outputs."""
import torch

import gnf_b200 as G
from helpers import load_golden, rel_err, rel_l2

def build_model(spec, device):
    return G.build_from_spec(spec, device)


def set_modes(model, mode):
    for c in model.getConditioners():
        for k, v in (mode or {}).items():
            setattr(c, k, v)


def run_case(name, device, ll_tol=1e-4, grad_tol=1e-3):
    c = load_golden(name)
    spec = c["spec"]
    model = build_model(spec, device)
    missing = model.load_state_dict({k: v.to(device) for k, v in c["sd"].items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    set_modes(model, c["mode"])
    noise = [n.to(device).contiguous() for n in c["noise"]]
    if noise:
        per = len(noise) // spec["nb_flow"]
        for k, cond in enumerate(model.getConditioners()):
            cond._replay_noise = tuple(noise[k * per:(k + 1) * per])
    x = c["x"].to(device)
    model.zero_grad()
    z, jac = model(x)
    loss = model.loss(z, jac)
    ll = model.z_log_density(z) + jac
    loss.backward()
    report = {"z": rel_err(z.detach().cpu(), c["z"]), "logdet": rel_err(jac.detach().cpu(), c["logdet"]),
              "ll": rel_err(ll.detach().cpu(), c["ll"]),
              "loss": abs(float(loss) - float(c["loss"])) / max(abs(float(c["loss"])), 1e-6)}
    params = dict(model.named_parameters())
    for k, g in c["grads"].items():
        assert params[k].grad is not None, f"{name}: no gradient for {k}"
        report["grad." + k] = rel_l2(params[k].grad.detach().cpu(), g)
    bad = {k: v for k, v in report.items() if not (v < (grad_tol if k.startswith("grad.") else ll_tol))}
    # tiny-magnitude z entries inflate the elementwise relative error; accept them on an absolute bar
    if "z" in bad and float((z.detach().cpu() - c["z"]).abs().max()) < 1e-5:
        bad.pop("z")
    if "logdet" in bad and float((jac.detach().cpu() - c["logdet"]).abs().max()) < 1e-5:
        bad.pop("logdet")
    assert not bad, f"{name}: out of tolerance {bad}\nfull report {report}"
    return report
