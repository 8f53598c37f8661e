"""Golden state trajectories of the reference's DAG dual-ascent control logic (DAGConditioner.step / update_dual_param /
post_process / constrainA / depth, models/Conditionners/DAGConditioner.py:76-92,171-174,196-293), recorded by running the
UNMODIFIED reference in the build container (needs /root/reference):

    python tests/golden/make_dag_control.py      ->  tests/golden/dag_control.npz

Shims that do not touch the arithmetic: oracle/UMNN.py for the absent UMNN pip package (imported by the reference's package
__init__, unused here) and networkx.from_numpy_matrix = from_numpy_array (networkx >= 3).  The reference's prints are muted.
"""
import contextlib
import io
import json
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

from models.Conditionners import DAGConditioner  # noqa: E402


def state(c):
    return dict(lambd=float(c.lambd), c=float(c.c), dag_const=float(c.dag_const), l1_weight=float(c.l1_weight),
                exponent=int(c.exponent), prev_trace=float(c.prev_trace), alpha=float(c.alpha), no_update=int(c.no_update),
                stoch_gate=bool(c.stoch_gate), noise_gate=bool(c.noise_gate), s_thresh=bool(c.s_thresh), h_thresh=float(c.h_thresh),
                is_invertible=bool(c.is_invertible), requires_grad=bool(c.A.requires_grad))


def tri(d, v=1.):
    return torch.tril(torch.ones(d, d), -1) * v


def scripts():
    g = torch.Generator().manual_seed(7)
    d = 5
    dense = torch.ones(d, d) * 1.5 + torch.randn(d, d, generator=g) * .02
    dense *= 1. - torch.eye(d)
    mixed = torch.randn(d, d, generator=g)
    small_cycle = tri(d, 1.2)
    small_cycle[0, 4] = .15                        # weak back edge: post_process must raise its threshold past soft(0.15) to cut it
    s = []
    # 0: the ordinary augmented-Lagrangian schedule on the default dense init: epochs without update, dual updates, c growth
    ops = [["step", [0, 10.]], ["step", [3, 10.]], ["step", [10, 1e9]], ["step", [20, 1e9]], ["step", [30, 1e-9]], ["step", [40, 1e9]],
           ["update_dual_param", []], ["update_dual_param", []]]
    s.append(dict(d=d, l1=.5, nb_epoch_update=10, A0=dense, ops=ops))
    # 1: constrainA with a threshold > 0 on mixed magnitudes (a11), then the exponent cut-back of step() when the trace > 50
    ops = [["constrainA", [.6]], ["constrainA", [.0001]], ["setA", dense * 3.], ["step", [1, 1.]], ["step", [2, 1.]], ["step", [7, 1.]]]
    s.append(dict(d=d, l1=0., nb_epoch_update=10, A0=mixed, ops=ops))
    # 2: an acyclic A: trace == 0 -> exponent growth -> post-processing succeeds -> dag_const = 0 -> 'no cycle' branch, depth
    ops = [["update_dual_param", []], ["update_dual_param", []], ["depth", []], ["step", [10, 1e9]]]
    s.append(dict(d=d, l1=.1, nb_epoch_update=10, A0=tri(d, 1.1), ops=ops))
    # 3: post_process threshold search on a graph with one weak back edge, and with an explicit threshold
    ops = [["post_process", []], ["depth", []]]
    s.append(dict(d=d, l1=0., nb_epoch_update=10, A0=small_cycle, ops=ops))
    ops = [["post_process", [.5]], ["depth", []]]
    s.append(dict(d=d, l1=0., nb_epoch_update=10, A0=small_cycle, ops=ops))
    # 5: dag_const forced to 0 while A still has cycles: the 'bad news' branch re-arms the constraint
    ops = [["set", ["stoch_gate", False]], ["set", ["s_thresh", False]], ["setdual", ["dag_const", 0.]], ["update_dual_param", []],
           ["update_dual_param", []]]
    s.append(dict(d=d, l1=0., nb_epoch_update=10, A0=dense, ops=ops))
    # 6: the no_update counter: eleven refused updates, then the forced one (d = 7: another exponent)
    ops = [["step", [10 * (k + 1), 1e-9]] for k in range(12)]
    s.append(dict(d=7, l1=0., nb_epoch_update=10, A0=None, ops=ops))
    # 7: exponent cut back to 3 by step() on a heavy A, then an acyclic A: the exponent grows by 50 before post-processing
    ops = [["step", [1, 1.]], ["setA", tri(d, .9)], ["update_dual_param", []], ["update_dual_param", []], ["depth", []]]
    s.append(dict(d=d, l1=0., nb_epoch_update=10, A0=dense * 3., ops=ops))
    return s


def main():
    out, meta = {}, []
    for si, sc in enumerate(scripts()):
        d = sc["d"]
        torch.manual_seed(100 + si)
        with contextlib.redirect_stdout(io.StringIO()):
            c = DAGConditioner(d, [8], 2, l1=sc["l1"], nb_epoch_update=sc["nb_epoch_update"], hot_encoding=True)
        if sc["A0"] is not None:
            with torch.no_grad():
                c.A.copy_(sc["A0"])
            c.prev_trace = c.get_power_trace().detach()
        c.A.grad = torch.zeros_like(c.A)                      # step() prints statistics of A.grad
        out[f"s{si}_A0"] = c.A.detach().numpy().copy()
        out[f"s{si}_prev_trace0"] = np.float32(float(c.prev_trace))
        states, ops = [], []
        for k, (op, args) in enumerate(sc["ops"]):
            with contextlib.redirect_stdout(io.StringIO()):
                extra = {}
                if op == "step":
                    c.step(args[0], torch.tensor(args[1]))
                elif op == "update_dual_param":
                    c.update_dual_param()
                elif op == "post_process":
                    with torch.no_grad():
                        c.post_process(*args)
                elif op == "constrainA":
                    with torch.no_grad():
                        c.constrainA(*args)
                elif op == "setA":
                    out[f"s{si}_setA{k}"] = args.numpy().copy()
                    with torch.no_grad():
                        c.A.data.copy_(args)
                    args = []
                elif op == "set":
                    setattr(c, args[0], args[1])
                elif op == "setdual":
                    getattr(c, args[0]).fill_(args[1])
                elif op == "depth":
                    extra["depth"] = int(c.depth())
                if c.A.requires_grad and c.A.grad is None:
                    c.A.grad = torch.zeros_like(c.A)
            st = state(c)
            st.update(extra)
            states.append(st)
            ops.append([op, args])
            out[f"s{si}_A{k + 1}"] = c.A.detach().numpy().copy()
        meta.append(dict(d=d, l1=sc["l1"], nb_epoch_update=sc["nb_epoch_update"], ops=ops, states=states))
    out["scripts"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, "dag_control.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for si, m in enumerate(meta):
        print(si, [(o[0], s["exponent"], round(s["lambd"], 6), s["c"], s["dag_const"], s["no_update"], s["is_invertible"], s["stoch_gate"]) for o, s in zip(m["ops"], m["states"])])


if __name__ == "__main__":
    main()
