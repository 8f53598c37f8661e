#!/usr/bin/env python
"""Command-line driver with the reference's flags (UCIExperiments.py:222-288) on the B200 backend.

    python examples/uci_experiments.py -load_config power-mono-DAG -config_file /path/to/UCIExperimentsConfigurations.yml -data power.npz
    python examples/uci_experiments.py -dataset synthetic -dim 6 -conditioner DAG -normalizer monotonic -emb_net 60 60 60 30 \\
           -int_net 100 100 100 -b_size 2500 -nb_epoch 3

-data: an .npz with arrays trn / val / tst [N, d] (the UCI loaders need h5py / the raw files, which this image does not have);
-dataset synthetic draws standardised Gaussian mixtures of dimension -dim.
"""
import argparse
import os
import sys
from datetime import datetime

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnf_b200 as G  # noqa: E402
from gnf_b200 import experiments as E  # noqa: E402


def synthetic(dim, n, seed=0):
    rng = np.random.RandomState(seed)
    mix = rng.randn(4, dim) * 2
    x = mix[rng.randint(0, 4, n)] + rng.randn(n, dim) * (.5 + rng.rand(dim))
    x = (x - x.mean(0)) / x.std(0)
    return x.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-load_config", default=None, type=str)
    ap.add_argument("-config_file", default="UCIExperimentsConfigurations.yml")
    ap.add_argument("-dataset", default="synthetic")
    ap.add_argument("-data", default=None, help=".npz with trn / val / tst")
    ap.add_argument("-dim", type=int, default=6)
    ap.add_argument("-folder", default="")
    ap.add_argument("-nb_flow", type=int, default=1)
    ap.add_argument("-weight_decay", default=1e-5, type=float)
    ap.add_argument("-learning_rate", default=1e-3, type=float)
    ap.add_argument("-nb_epoch", default=10000, type=int)
    ap.add_argument("-b_size", default=100, type=int)
    ap.add_argument("-conditioner", default="DAG", choices=sorted(E.COND_TYPES))
    ap.add_argument("-emb_net", default=[100, 100, 100, 10], nargs="+", type=int)
    ap.add_argument("-nb_steps_dual", default=100, type=int)
    ap.add_argument("-l1", default=.2, type=float)
    ap.add_argument("-gumble_T", default=1., type=float)
    ap.add_argument("-normalizer", default="affine", choices=sorted(E.NORM_TYPES))
    ap.add_argument("-int_net", default=[100, 100, 100, 100], nargs="+", type=int)
    ap.add_argument("-nb_steps", default=20, type=int)
    ap.add_argument("-solver", default="CC", choices=["CC", "CCParallel"])
    ap.add_argument("-no_graph", action="store_true")
    args = ap.parse_args()
    cfg = dict(E.DRIVER_DEFAULTS)
    cfg.update({k: v for k, v in vars(args).items() if k in cfg})
    if args.load_config is not None:
        preset = E.load_preset(args.config_file, args.load_config)
        nb_epoch = cfg["nb_epoch"]
        cfg.update(preset)
        if "-nb_epoch" in sys.argv:
            cfg["nb_epoch"] = nb_epoch
    if args.data:
        f = np.load(args.data)
        trn, val, tst = f["trn"], f["val"], f["tst"]
    else:
        trn, val, tst = synthetic(args.dim, 20 * cfg["b_size"], 0), synthetic(args.dim, 4 * cfg["b_size"], 1), synthetic(args.dim, 4 * cfg["b_size"], 2)
    name = args.load_config or cfg["dataset"] or "synthetic"
    path = args.folder or os.path.join("UCIExperiments", name, datetime.now().strftime("%m_%d_%Y_%H_%M_%S"))
    os.makedirs(path, exist_ok=True)
    E.train_uci(trn, val, tst, cfg, path=path, use_graph=not args.no_graph)


if __name__ == "__main__":
    main()
