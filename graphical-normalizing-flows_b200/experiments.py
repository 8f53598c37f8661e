"""Host-side experiment plumbing around the hot path: the reference drivers' training loop shape
(UCIExperiments.py:58-220), its YAML presets (UCIExperimentsConfigurations.yml), bits-per-pixel (ImageExperiments.py:33-37) and
checkpoints written under nn.DataParallel (`module.`-prefixed keys, ImageExperiments.py:168,177,252).  Pure Python; every
tensor operation of the loop goes through the flow's CUDA path."""
import contextlib
import math
import os
import time

import numpy as np
import torch

from .conditioners import AutoregressiveConditioner, CouplingConditioner, DAGConditioner
from .flow import buildFCNormalizingFlow
from .graphs import GraphedTrainStep
from .normalizers import AffineNormalizer, MonotonicNormalizer
from .optim import FusedAdam
from . import dist as gdist

COND_TYPES = {"DAG": DAGConditioner, "Coupling": CouplingConditioner, "Autoregressive": AutoregressiveConditioner}
NORM_TYPES = {"affine": AffineNormalizer, "monotonic": MonotonicNormalizer}
# argparse defaults of the reference driver (UCIExperiments.py:226-256); a preset overrides them key by key
DRIVER_DEFAULTS = dict(dataset=None, nb_flow=1, weight_decay=1e-5, learning_rate=1e-3, nb_epoch=10000, b_size=100, conditioner="DAG",
                       emb_net=[100, 100, 100, 10], nb_steps_dual=100, l1=.2, gumble_T=1., normalizer="affine",
                       int_net=[100, 100, 100, 100], nb_steps=20, solver="CC")


def compute_bpp(ll, x, alpha=1e-6):
    """Bits per pixel of logit-transformed images (ImageExperiments.py:33-37): ll [B] log-likelihood of x [B, d] (logit space)."""
    d = x.shape[1]
    sig = torch.sigmoid(x)
    return -ll / (d * math.log(2)) - math.log2(1 - 2 * alpha) + 8 + (torch.log2(sig) + torch.log2(1. - sig)).sum(1) / d


def strip_data_parallel_prefix(state_dict):
    """Checkpoints saved from `nn.DataParallel(model).state_dict()` carry a `module.` prefix on every key."""
    if state_dict and all(k.startswith("module.") for k in state_dict):
        return {k[len("module."):]: v for k, v in state_dict.items()}
    return state_dict


def load_checkpoint(model, path_or_state, strict=True, map_location=None):
    """model.load_state_dict for a reference checkpoint file (or an already loaded dict), DataParallel-prefixed or not."""
    sd = path_or_state
    if not isinstance(sd, dict):
        sd = torch.load(path_or_state, map_location=map_location or "cpu")
    return model.load_state_dict(strip_data_parallel_prefix(sd), strict=strict)


def load_preset(yaml_path, name):
    """One named experiment of a reference-format YAML file (e.g. the reference's own UCIExperimentsConfigurations.yml) merged
    over the driver's argparse defaults.  The reference patches PyYAML's float resolver so that `1e-5` parses as a float; here
    numeric strings are converted after loading."""
    import yaml
    with open(yaml_path) as f:
        allcfg = yaml.safe_load(f)
    if name not in allcfg:
        raise KeyError(f"{name} is not in {yaml_path}: {sorted(allcfg)[:8]}...")
    cfg = dict(DRIVER_DEFAULTS)
    for k, v in allcfg[name].items():
        if isinstance(v, str):
            try:
                v = float(v)
            except ValueError:
                pass
        cfg[k] = v
    return cfg


def build_uci_flow(dim, cfg):
    """buildFCNormalizingFlow with the argument dictionaries the reference driver assembles (UCIExperiments.py:81-95)."""
    ctype, ntype = COND_TYPES[cfg["conditioner"]], NORM_TYPES[cfg["normalizer"]]
    emb = list(cfg["emb_net"])
    cargs = {"in_size": dim, "hidden": emb[:-1], "out_size": emb[-1]}
    if ctype is DAGConditioner:
        cargs.update(l1=cfg["l1"], gumble_T=.5, nb_epoch_update=cfg["nb_steps_dual"], hot_encoding=True)
    nargs = {}
    if ntype is MonotonicNormalizer:
        nargs = {"integrand_net": list(cfg["int_net"]), "cond_size": emb[-1], "nb_steps": cfg["nb_steps"], "solver": cfg["solver"]}
    return buildFCNormalizingFlow(cfg["nb_flow"], ctype, cargs, ntype, nargs)


def _batches(x, batch_size, shuffle, generator=None):
    idx = torch.randperm(x.shape[0], device=x.device, generator=generator) if shuffle else torch.arange(x.shape[0], device=x.device)
    for s in range(0, x.shape[0] - batch_size + 1, batch_size):          # whole batches only: the captured step has a fixed shape
        yield x[idx[s:s + batch_size]]


@torch.no_grad()
def mean_log_likelihood(model, x, batch_size):
    tot, n = 0., 0
    for cur in _batches(x, min(batch_size, x.shape[0]), False):
        ll, _ = model.compute_ll(cur)
        tot += float(ll.mean())
        n += 1
    return tot / max(n, 1)


def train_uci(trn, val, tst, cfg, path=None, device="cuda", nb_epoch=None, use_graph=True, log=print, seed=None):
    """The reference's UCI training loop on tensors trn / val / tst [N, d]: per epoch constrainA(0), shuffled fixed-size batches
    with S = nb_steps + U{0..9} quadrature steps for Monotonic flows, Adam(lr, weight_decay), model.step(epoch, mean loss),
    validation with nb_steps + 20, `best_model.pt` when DAGness < 1e-20 and the validation loss improves, model / ADAM checkpoints.
    The training step replays as CUDA graphs (one per S) unless use_graph is False.  Returns the per-epoch history."""
    if seed is not None:
        torch.manual_seed(seed)
    dev = torch.device(device)
    trn, val, tst = (torch.as_tensor(a, dtype=torch.float32).to(dev) for a in (trn, val, tst))
    dim = trn.shape[1]
    model = build_uci_flow(dim, cfg).to(dev)
    gdist.broadcast_parameters(model)
    opt = FusedAdam(model.parameters(), lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])
    bucket = gdist.GradBucket(model.parameters(), overlap=False)
    mono = NORM_TYPES[cfg["normalizer"]] is MonotonicNormalizer
    is_dag = COND_TYPES[cfg["conditioner"]] is DAGConditioner
    bsz, S0 = int(cfg["b_size"]), int(cfg["nb_steps"])
    graphs, history, best, stepped = {}, [], float("inf"), False
    rnd = np.random.RandomState(0 if seed is None else seed)

    def eager_step(x):
        bucket.begin_step()
        z, jac = model(x)
        loss = model.loss(z, jac)
        loss.backward()
        bucket.finish_step()
        opt.step()
        return loss.detach()

    # the whole loop -- eager steps, captures, replays, validation -- runs on one side stream (see graphs._capture)
    train_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
    if train_stream is not None:
        train_stream.wait_stream(torch.cuda.current_stream(dev))
    with (torch.cuda.stream(train_stream) if train_stream is not None else contextlib.nullcontext()):
        for epoch in range(int(nb_epoch if nb_epoch is not None else cfg["nb_epoch"])):
            t0 = time.perf_counter()
            if is_dag:
                with torch.no_grad():
                    for c in model.getConditioners():
                        c.constrainA(zero_threshold=0.)
            tot, count = torch.zeros((), device=dev), 0
            for cur in _batches(trn, bsz, True):
                S = S0 + int(rnd.randint(0, 10)) if mono else S0
                if mono:
                    for nrm in model.getNormalizers():
                        nrm.nb_steps = S
                if use_graph and stepped:
                    if S not in graphs:
                        graphs[S] = GraphedTrainStep(model, opt, bucket, cur, allreduce=True, warmup=2, stream=train_stream)
                    loss = graphs[S](cur)
                else:
                    # also the very first batch with graphs on: autograd binds every parameter's gradient accumulator to the stream of
                    # its first backward; that must be the ordinary stream, not the side stream of one capture's warm-up (a later
                    # capture would otherwise be invalidated by a dependency on that stream)
                    loss = eager_step(cur)
                stepped = True
                tot += loss
                count += 1
            tot = tot / max(count, 1)
            if not math.isfinite(float(tot)):
                if path:
                    torch.save(model.state_dict(), os.path.join(path, "NANmodel.pt"))
                raise FloatingPointError("NaN / inf in the training loss")
            model.step(epoch, tot)                      # dual ascent; a change of the DAG state makes the graphs recapture
            if mono:
                for nrm in model.getNormalizers():
                    nrm.nb_steps = S0 + 20
            ll_val = mean_log_likelihood(model, val, bsz)
            dagness = max(model.DAGness()) if is_dag else 0.
            rec = dict(epoch=epoch, train_loss=float(tot), valid_ll=ll_val, dagness=float(dagness), seconds=time.perf_counter() - t0)
            if dagness < 1e-20 and -ll_val < best:
                best = -ll_val
                rec["test_ll"] = mean_log_likelihood(model, tst, bsz)
                if path:
                    torch.save(model.state_dict(), os.path.join(path, "best_model.pt"))
            history.append(rec)
            log("epoch: {epoch:d} - Train loss: {train_loss:4f} - Valid log-likelihood: {valid_ll:4f} - <<DAGness>>: {dagness:4f} - "
                "Elapsed time per epoch {seconds:4f} (seconds)".format(**rec))
            if path:
                torch.save(model.state_dict(), os.path.join(path, "model_%d.pt" % epoch))
                torch.save(opt.state_dict(), os.path.join(path, "ADAM_%d.pt" % epoch))
    if train_stream is not None:
        torch.cuda.current_stream(dev).wait_stream(train_stream)
    if path:
        torch.save(model.state_dict(), os.path.join(path, "model.pt"))
        torch.save(opt.state_dict(), os.path.join(path, "ADAM.pt"))
    return model, history
