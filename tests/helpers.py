"""Shared test helpers: golden-fixture loading and error metrics."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith(("power_trace.npz", "dag_control.npz", "image_cnn_flow.npz")))


def load_golden(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = {"spec": ast.literal_eval(str(f["spec"])), "mode": ast.literal_eval(str(f["mode"]))}
    case["x"] = torch.from_numpy(f["x"])
    for k in ("z", "logdet", "ll", "loss"):
        case[k] = torch.from_numpy(f[k])
    case["sd"] = {k[3:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("sd.")}
    case["grads"] = {k[5:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("grad.")}
    noise = sorted((int(k.split(".")[1]), k) for k in f.files if k.startswith("noise."))
    case["noise"] = [torch.from_numpy(f[k]) for _, k in noise]
    return case


def noises_per_step(case):
    """Split the flat replay queue into per-flow-step tuples."""
    n = case["noise"]
    nb = case["spec"]["nb_flow"]
    if not n:
        return None
    per = len(n) // nb
    return [tuple(n[i * per:(i + 1) * per]) for i in range(nb)]


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) elementwise -> max."""
    a, b = a.double(), b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-6)).max())


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    nb = float(b.norm())
    if nb == 0.:
        return float(a.norm())
    return float((a - b).norm()) / nb
