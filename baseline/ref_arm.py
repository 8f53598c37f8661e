"""The reference arm of bench.py: the UNMODIFIED reference classes (baseline/_ref/models, see install_ref.py) driven
through the reference's own public API and stock code path on the host CPU -- buildFCNormalizingFlow(...), model(x),
model.loss(z, jac), loss.backward(), torch.optim.Adam; evaluation = model(x) + model.z_log_density(z) (UCIExperiments.py:96-162).
None of this repository's kernels, models or engine is on that path; `oracle/UMNN.py` stands in for the absent UMNN==1.0
pip dependency (so the arm is "reference + restated UMNN")."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "models", "NormalizingFlowFactories.py"))


def _import_reference():
    for p in (os.path.join(ROOT, "oracle"), REF):          # oracle/ provides the `UMNN` module the reference imports
        if p not in sys.path:
            sys.path.insert(0, p)
    import networkx as nx
    if not hasattr(nx, "from_numpy_matrix"):                # networkx >= 3 (only depth / post_process / update_dual_param use it)
        nx.from_numpy_matrix = nx.from_numpy_array
    import models as R                                       # baseline/_ref/models
    assert os.path.realpath(os.path.dirname(R.__file__)).startswith(os.path.realpath(REF)), R.__file__
    return R


def build_reference_model(spec, seed=0):
    import torch
    R = _import_reference()
    from models.NormalizingFlowFactories import MNIST_A_prior
    cond = {"DAG": R.DAGConditioner, "Autoregressive": R.AutoregressiveConditioner, "Coupling": R.CouplingConditioner}[spec["cond"]]
    torch.manual_seed(seed)
    cargs = {"in_size": spec["d"], "hidden": list(spec["hidden"]), "out_size": spec["out"]}
    if spec["cond"] == "DAG":
        cargs.update(l1=spec.get("l1", 0.), gumble_T=spec.get("gumble_T", 1.), nb_epoch_update=10,
                     hot_encoding=spec.get("hot_encoding", False))
        if spec.get("A_prior") == "mnist":
            cargs["A_prior"] = MNIST_A_prior(int(round(spec["d"] ** .5)), 2)
    if spec["norm"] == "monotonic":
        ntype = R.MonotonicNormalizer
        nargs = {"integrand_net": list(spec["int_net"]), "cond_size": spec["out"], "nb_steps": spec["nb_steps"],
                 "solver": spec.get("solver", "CC")}
    else:
        ntype, nargs = R.AffineNormalizer, {}
    return R.buildFCNormalizingFlow(spec["nb_flow"], cond, cargs, ntype, nargs)


def step_fn(spec, B, mode, S, lr, wd):
    """One training (zero_grad -> forward -> loss -> backward -> Adam.step) or evaluation (compute_ll) step of the reference."""
    import torch
    model = build_reference_model(spec)
    for n in model.getNormalizers():
        if hasattr(n, "nb_steps"):
            n.nb_steps = S
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    g = torch.Generator().manual_seed(0)
    d = spec["d"]

    def train_step():
        x = torch.randn(B, d, generator=g)
        opt.zero_grad()
        z, jac = model(x)
        loss = model.loss(z, jac)
        loss.backward()
        opt.step()
        return float(loss.detach())

    def eval_step():
        x = torch.randn(B, d, generator=g)
        with torch.no_grad():                                # UCIExperiments.py:152-158 (this snapshot's model has no compute_ll)
            z, jac = model(x)
            ll = model.z_log_density(z) + jac
        return float(ll.mean())

    return train_step if mode == "train" else eval_step
