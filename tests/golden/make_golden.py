"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ``models`` package is imported from /root/reference with two shims that do
not touch the hot path's arithmetic: ``oracle/UMNN.py`` stands in for the absent
``UMNN==1.0`` pip dependency, and ``networkx.from_numpy_matrix`` is aliased to
``from_numpy_array`` (networkx >= 3).  The stochastic gate's ``torch.rand`` draws
(DAGConditioner.py:99-100) are replayed from recorded tensors so that the same noise can
be fed to the CUDA path.  Outputs: tests/golden/*.npz (small fixtures, committed).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import networkx as nx  # noqa: E402

if not hasattr(nx, "from_numpy_matrix"):
    nx.from_numpy_matrix = nx.from_numpy_array

from models.Conditionners import DAGConditioner, AutoregressiveConditioner, CouplingConditioner  # noqa: E402
from models.Normalizers import AffineNormalizer, MonotonicNormalizer  # noqa: E402
from models.NormalizingFlowFactories import buildFCNormalizingFlow  # noqa: E402

COND = {"DAG": DAGConditioner, "Autoregressive": AutoregressiveConditioner, "Coupling": CouplingConditioner}


class ReplayRand:
    """Context manager: torch.rand / torch.randn inside the reference return queued tensors."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def pop(shape, *a, **k):
            t = self.queue.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        torch.rand = pop
        torch.randn = pop
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def build_reference(spec, seed):
    torch.manual_seed(seed)
    cargs = {"in_size": spec["d"], "hidden": list(spec["hidden"]), "out_size": spec["out"]}
    if spec["cond"] == "DAG":
        cargs.update(l1=spec.get("l1", 0.), gumble_T=spec.get("gumble_T", 1.), nb_epoch_update=10,
                     hot_encoding=spec.get("hot_encoding", False))
    if spec["norm"] == "monotonic":
        ntype = MonotonicNormalizer
        nargs = {"integrand_net": list(spec["int_net"]), "cond_size": spec["out"], "nb_steps": spec["nb_steps"],
                 "solver": spec.get("solver", "CC")}
    else:
        ntype, nargs = AffineNormalizer, {}
    return buildFCNormalizingFlow(spec["nb_flow"], COND[spec["cond"]], cargs, ntype, nargs)


def run_reference(model, spec, x, mode=None, noise_queue=(), exponents=None):
    conds = model.getConditioners()
    if spec["cond"] == "DAG":
        for k, c in enumerate(conds):
            for key, val in (mode or {}).items():
                setattr(c, key, val)
            if exponents is not None:
                c.exponent = exponents[k]
    model.zero_grad()
    with ReplayRand(noise_queue):
        z, jac = model(x)
    loss = model.loss(z, jac)
    ll = model.z_log_density(z) + jac
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return z.detach(), jac.detach(), ll.detach(), loss.detach(), grads


def make_noise(spec, B, mode, gen):
    d = spec["d"]
    m = dict(s_thresh=True, h_thresh=0., stoch_gate=True, noise_gate=False)
    m.update(mode or {})
    if spec["cond"] != "DAG" or not (m["h_thresh"] > 0 or m["s_thresh"]):
        return []
    q = []
    for _ in range(spec["nb_flow"]):
        if m["stoch_gate"]:
            q += [torch.rand(B, d, d, generator=gen), torch.rand(B, d, d, generator=gen)]
        elif m["noise_gate"]:
            q += [torch.randn(B, d, d, generator=gen)]
    return q


CASES = {
    "toy_dag_affine": dict(spec=dict(nb_flow=3, d=2, cond="DAG", hidden=[16, 16], out=16, hot_encoding=True,
                                     gumble_T=.5, l1=1., norm="affine"), B=8),
    "power_dag_mono": dict(spec=dict(nb_flow=1, d=6, cond="DAG", hidden=[12, 12, 12], out=5, hot_encoding=True,
                                     gumble_T=.5, l1=0., norm="monotonic", int_net=[10, 10, 10], nb_steps=7,
                                     solver="CC"), B=5),
    "dag_mono_nohot": dict(spec=dict(nb_flow=2, d=4, cond="DAG", hidden=[9], out=3, hot_encoding=False,
                                     gumble_T=1., l1=.3, norm="monotonic", int_net=[7, 11], nb_steps=5,
                                     solver="CCParallel"), B=4),
    "ar_mono": dict(spec=dict(nb_flow=1, d=5, cond="Autoregressive", hidden=[15, 15], out=4, norm="monotonic",
                              int_net=[8, 8, 8], nb_steps=6, solver="CCParallel"), B=6),
    "ar_affine": dict(spec=dict(nb_flow=2, d=4, cond="Autoregressive", hidden=[12, 12, 12], out=2, norm="affine"),
                      B=7),
    "coupling_affine": dict(spec=dict(nb_flow=2, d=5, cond="Coupling", hidden=[9, 9], out=3, norm="affine"), B=6),
    "coupling_mono": dict(spec=dict(nb_flow=1, d=4, cond="Coupling", hidden=[6, 6], out=3, norm="monotonic",
                                    int_net=[5, 5, 5], nb_steps=4, solver="CC"), B=3),
    # deterministic / thresholded / post-processed gating branches (DAGConditioner.py:126-153)
    "dag_soft_det": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=True, gumble_T=.5,
                                   l1=.5, norm="affine"), B=4, mode=dict(stoch_gate=False)),
    "dag_hard_stoch": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=True,
                                     gumble_T=.5, l1=.5, norm="affine"), B=4, mode=dict(h_thresh=.3), scaleA=.35),
    "dag_hard_det": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=True, gumble_T=.5,
                                   l1=.5, norm="affine"), B=4, mode=dict(h_thresh=.3, stoch_gate=False),
                         scaleA=.35),
    "dag_hard_a2": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=False, gumble_T=.5,
                                  l1=.5, norm="affine"), B=4,
                        mode=dict(h_thresh=.1, s_thresh=False, stoch_gate=False), scaleA=.35),
    "dag_noise_gate": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=True,
                                     gumble_T=.5, l1=.5, norm="affine"), B=4,
                           mode=dict(stoch_gate=False, noise_gate=True), scaleA=.5),
    "dag_raw_A": dict(spec=dict(nb_flow=1, d=5, cond="DAG", hidden=[8, 8], out=2, hot_encoding=True, gumble_T=.5,
                                l1=0., norm="affine"), B=4, mode=dict(stoch_gate=False, s_thresh=False),
                      binaryA=True),
}


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, case in CASES.items():
        spec, B = case["spec"], case["B"]
        gen = torch.Generator().manual_seed(1234)
        model = build_reference(spec, seed=7)
        with torch.no_grad():
            if "scaleA" in case:
                for c in model.getConditioners():
                    c.A.mul_(case["scaleA"] * torch.rand(c.A.shape, generator=gen) * 2)
            if case.get("binaryA"):
                for c in model.getConditioners():
                    c.A.copy_(torch.tril((torch.rand(c.A.shape, generator=gen) > .4).float(), -1))
        x = torch.randn(B, spec["d"], generator=gen)
        noise = make_noise(spec, B, case.get("mode"), gen)
        exps = None
        z, jac, ll, loss, grads = run_reference(model, spec, x, case.get("mode"), noise, exps)
        out = {"x": x.numpy(), "z": z.numpy(), "logdet": jac.numpy(), "ll": ll.numpy(), "loss": loss.numpy()}
        for i, n in enumerate(noise):
            out[f"noise.{i}"] = n.numpy()
        for k, v in model.state_dict().items():
            out[f"sd.{k}"] = v.detach().numpy()
        for k, v in grads.items():
            out[f"grad.{k}"] = v.numpy()
        out["spec"] = np.array(repr(spec))
        out["mode"] = np.array(repr(case.get("mode")))
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(f"{name}: loss={float(loss):.6g} ll[0]={float(ll[0]):.6g} "
              f"{sum(v.size for v in out.values() if hasattr(v, 'size'))} values")

    # power-trace known values straight from the reference method (DAGConditioner.py:176-194)
    pt = {}
    gen = torch.Generator().manual_seed(99)
    for d, p, scale in [(2, 2, 1.5), (6, 6, 1.5), (21, 21, .6), (63, 13, 1.5), (63, 13, .2), (40, 3, 1.), (100, 50, .3)]:
        torch.manual_seed(d)
        c = DAGConditioner(d, [4], 2)
        with torch.no_grad():
            c.A.copy_((scale + .02 * torch.randn(d, d, generator=gen)) * (1 - torch.eye(d)))
        c.exponent = p
        c.A.grad = None
        t = c.get_power_trace()
        t.backward()
        key = f"d{d}_p{p}_s{scale}"
        pt[f"{key}.A"] = c.A.detach().numpy().copy()
        pt[f"{key}.t"] = t.detach().numpy()
        pt[f"{key}.dA"] = c.A.grad.numpy().copy()
        pt[f"{key}.meta"] = np.array([d, p, 1. / d])
        print(key, float(t))
    np.savez_compressed(os.path.join(HERE, "power_trace.npz"), **pt)


if __name__ == "__main__":
    main()
