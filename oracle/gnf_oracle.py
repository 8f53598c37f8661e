"""CPU oracle for the Graphical-Normalizing-Flows hot path (TEST INFRASTRUCTURE ONLY).

A functional, torch-CPU fp32 restatement of the reference's density-evaluation / training
path.  Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  It is *pinned* against the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference ``models`` package in the build container (with
``oracle/UMNN.py`` standing in for the absent ``UMNN==1.0`` pip dependency), runs it on
seeded inputs and commits the input/output vectors under ``tests/golden/``;
``tests/test_oracle.py`` checks this oracle against those vectors.  The UMNN integral
itself has no reference-side fixture -> "parity unpinned" at that one boundary (see
``oracle/UMNN.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path never does and fails
loudly when its CUDA library is missing.

State is passed as a flat ``state_dict`` with the reference's own key names
(``steps.{k}.conditioner.A`` ...), plus a ``spec`` dict describing the architecture:

    spec = dict(nb_flow=1, d=6, cond="DAG"|"Autoregressive"|"Coupling", hidden=[..], out=30,
                hot_encoding=True, gumble_T=.5, norm="affine"|"monotonic",
                int_net=[..], nb_steps=20)
and an optional per-step ``mode`` dict for the DAG conditioner's mutable Python attributes
(``s_thresh, h_thresh, stoch_gate, noise_gate, exponent``; defaults = the constructor's).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS_GATE = 1e-6


# ----------------------------------------------------------------------------------------------
# Clenshaw-Curtis quadrature (UMNN==1.0; call site MonotonicNormalizer.py:58-63)
# ----------------------------------------------------------------------------------------------
def cc_weights_nodes(nb_steps):
    """float64 numpy CC weights w_k and nodes cos(k*pi/S), cast to fp32 (App. B)."""
    S = nb_steps
    k = np.arange(S + 1, dtype=np.float64)
    lam = np.cos(np.outer(k, k) * math.pi / S)
    lam[:, 0] = .5
    lam[:, -1] = .5 * lam[:, -1]
    lam = lam * 2 / S
    W = np.zeros(S + 1, dtype=np.float64)
    even = np.arange(0, S + 1, 2)
    W[even] = 2. / (1. - even.astype(np.float64) ** 2)
    W[0] = 1.
    w = lam.T @ W
    return torch.tensor(w).float(), torch.tensor(np.cos(k * math.pi / S)).float()


def mlp(x, sd, prefix, n_layers, idx_stride=2):
    """Linear/ReLU stack without final activation (DAGConditioner.py:7-20,
    CouplingConditioner.py:6-18, MonotonicNormalizer.py:21-31)."""
    for l in range(n_layers):
        x = F.linear(x, sd[f"{prefix}.{idx_stride * l}.weight"], sd[f"{prefix}.{idx_stride * l}.bias"])
        if l < n_layers - 1:
            x = torch.relu(x)
    return x


def integrand(xv, hrow, sd, prefix, n_layers):
    """IntegrandNet.forward restated row-wise (MonotonicNormalizer.py:12-38).

    xv [N] scalar inputs, hrow [N, E] embeddings -> f [N] = ELU(MLP([x, h])) + 1.05."""
    inp = torch.cat((xv.unsqueeze(1), hrow), 1)
    y = mlp(inp, sd, prefix, n_layers).squeeze(1)
    return F.elu(y) + 1.05


def monotonic_normalizer(x, h, sd, prefix, n_layers, nb_steps):
    """MonotonicNormalizer.forward (MonotonicNormalizer.py:51-66) with the UMNN integral
    (App. B) written as explicit quadrature + a custom backward that follows UMNN's gradient
    convention (Leibniz rule for dx, quadrature of parameter / h gradients)."""
    B, d = x.shape
    E = h.shape[2]
    hr = h.reshape(B * d, E)
    xr = x.reshape(B * d)
    z_int = _UMNNRowIntegral.apply(xr, hr, sd, prefix, n_layers, nb_steps,
                                   *[sd[f"{prefix}.{2 * l}.{n}"] for l in range(n_layers) for n in ("weight", "bias")])
    z = z_int.view(B, d) + h[:, :, 0]
    jac = integrand(xr, hr, sd, prefix, n_layers).view(B, d)
    return z, jac


class _UMNNRowIntegral(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xr, hr, sd, prefix, n_layers, nb_steps, *params):
        w, t = cc_weights_nodes(nb_steps)
        with torch.no_grad():
            acc = torch.zeros_like(xr)
            for k in range(nb_steps + 1):                       # x0 = 0 (MonotonicNormalizer.py:52)
                Xk = xr * (t[k] + 1) / 2
                acc = acc + w[k] * integrand(Xk, hr, sd, prefix, n_layers)
            out = acc * xr / 2
        ctx.save_for_backward(xr, hr)
        ctx.meta = (sd, prefix, n_layers, nb_steps, params)
        return out

    @staticmethod
    def backward(ctx, g):
        xr, hr = ctx.saved_tensors
        sd, prefix, n_layers, nb_steps, params = ctx.meta
        w, t = cc_weights_nodes(nb_steps)
        g_params = [torch.zeros_like(p) for p in params]
        g_h = torch.zeros_like(hr)
        with torch.enable_grad():
            hr_ = hr.detach().requires_grad_(True)
            sd_ = dict(sd)
            leaf = []
            i = 0
            for l in range(n_layers):
                for n in ("weight", "bias"):
                    p = params[i].detach().requires_grad_(True)
                    sd_[f"{prefix}.{2 * l}.{n}"] = p
                    leaf.append(p)
                    i += 1
            for k in range(nb_steps + 1):
                Xk = (xr * (t[k] + 1) / 2).detach()
                f = integrand(Xk, hr_, sd_, prefix, n_layers)
                cot = (g * xr / 2 * w[k]).detach()
                grads = torch.autograd.grad(f, leaf + [hr_], cot)
                for a, b in zip(g_params, grads[:-1]):
                    a += b
                g_h += grads[-1]
        with torch.no_grad():
            gx = integrand(xr, hr, sd, prefix, n_layers) * g            # Leibniz
        return (gx, g_h, None, None, None, None, *g_params)


# ----------------------------------------------------------------------------------------------
# Normalizers
# ----------------------------------------------------------------------------------------------
def affine_normalizer(x, h):
    """AffineNormalizer.forward (AffineNormalizer.py:9-12); out-of-place clamp (the in-place
    write-back of the clamped values into h is a side effect the tests check separately)."""
    mu = h[:, :, 0].clamp(-5., 5.)
    sigma = torch.exp(h[:, :, 1].clamp(-5., 2.))
    return x * sigma + mu, sigma


def normal_log_density(z):
    """NormalLogDensity.forward (NormalizingFlowFactories.py:15-16); pi is an fp32 buffer."""
    pi = torch.tensor(math.pi)
    return -.5 * (torch.log(pi * 2) + z ** 2).sum(1)


# ----------------------------------------------------------------------------------------------
# DAG conditioner
# ----------------------------------------------------------------------------------------------
def soft_thresholded_A(A):
    """DAGConditioner.py:118-119."""
    return 2 * (torch.sigmoid(2 * (A ** 2)) - .5)


def hard_thresholded_A(A, s_thresh, h_thresh):
    """DAGConditioner.py:121-124."""
    if s_thresh:
        s = soft_thresholded_A(A)
        return s * (s > h_thresh).float()
    return A ** 2 * (A ** 2 > h_thresh).float()


def gumbel_gate(importance, u1, u2, T):
    """DAGConditioner.stochastic_gate, Gumbel branch (DAGConditioner.py:94-103) with the two
    uniform draws supplied by the caller (the replay hook the parity tests use)."""
    g1 = -torch.log(-torch.log(u1))
    g2 = -torch.log(-torch.log(u2))
    z1 = torch.exp((torch.log(importance + EPS_GATE) + g1) / T)
    z2 = torch.exp((torch.log(1 - importance + EPS_GATE) + g2) / T)
    return z1 / (z1 + z2)


DEFAULT_MODE = dict(s_thresh=True, h_thresh=0., stoch_gate=True, noise_gate=False)


def dag_masked_input(x, A, mode, T, noise):
    """The 7-way gating branch of DAGConditioner.forward (DAGConditioner.py:126-153).
    Returns e [B*d, d] with e[b*d+i, j] = x[b, j] * G[b, i, j]."""
    B, d = x.shape
    m = dict(DEFAULT_MODE)
    m.update(mode or {})
    xe = x.unsqueeze(1).expand(-1, d, -1)
    if m["h_thresh"] > 0:
        imp = hard_thresholded_A(A, m["s_thresh"], m["h_thresh"])
    elif m["s_thresh"]:
        imp = soft_thresholded_A(A)
    else:
        return (xe * A.unsqueeze(0)).reshape(B * d, d)
    impe = imp.unsqueeze(0).expand(B, -1, -1)
    if m["stoch_gate"]:
        u1, u2 = noise
        e = xe * gumbel_gate(impe, u1, u2, T)
    elif m["noise_gate"]:
        (n,) = noise                                            # DAGConditioner.py:114-116
        e = impe * (xe + n * torch.sqrt((1 - impe) ** 2))
    else:
        e = xe * impe
    return e.reshape(B * d, d)


def dag_conditioner(x, sd, prefix, spec, mode=None, noise=None):
    """DAGConditioner.forward (DAGConditioner.py:126-169) -> h [B, d, H]."""
    B, d = x.shape
    e = dag_masked_input(x, sd[f"{prefix}.A"], mode, spec.get("gumble_T", 1.), noise)
    if spec.get("hot_encoding", False):
        hot = torch.eye(d).unsqueeze(0).expand(B, -1, -1).reshape(-1, d)
        e = torch.cat((e, hot), 1)
    n_layers = len(spec["hidden"]) + 1
    return mlp(e, sd, f"{prefix}.embedding_net.net", n_layers).view(B, d, -1)


def matrix_power(Bm, p):
    """torch.matrix_power's multiplication order (binary decomposition; n<=3 special-cased)."""
    if p == 0:
        return torch.eye(Bm.shape[0])
    if p == 1:
        return Bm.clone()
    if p == 2:
        return Bm @ Bm
    if p == 3:
        return (Bm @ Bm) @ Bm
    z, res = None, None
    while p > 0:
        bit, p = p % 2, p // 2
        z = Bm if z is None else z @ z
        if bit:
            res = z if res is None else res @ z
    return res


def power_trace(A, alpha, exponent):
    """DAGConditioner.get_power_trace, non-Hutchinson branch (DAGConditioner.py:176-194)."""
    d = A.shape[0]
    a = min(1., float(alpha))
    Bm = torch.eye(d) + a * A ** 2
    return torch.diag(matrix_power(Bm, exponent)).sum() - d


def dag_loss(sd, prefix, exponent):
    """DAGConditioner.loss (DAGConditioner.py:268-271)."""
    A = sd[f"{prefix}.A"]
    t = power_trace(A, sd[f"{prefix}.alpha"], exponent)
    return sd[f"{prefix}.dag_const"] * (sd[f"{prefix}.lambd"] * t + sd[f"{prefix}.c"] / 2 * t ** 2) \
        + sd[f"{prefix}.l1_weight"] * A.abs().mean()


# ----------------------------------------------------------------------------------------------
# Autoregressive (MADE) and coupling conditioners
# ----------------------------------------------------------------------------------------------
def made_masks(nin, hidden_sizes, nout):
    """MADE.update_masks, natural ordering, non-random (AutoregressiveConditioner.py:70-106).
    Returned masks are [out_features, in_features] like MaskedLinear.mask."""
    L = len(hidden_sizes)
    m = {-1: np.arange(nin)}
    for l in range(L):
        m[l] = np.array([nin - 1 - (i % nin) for i in range(hidden_sizes[l])])
    masks = [m[l - 1][:, None] <= m[l][None, :] for l in range(L)]
    masks.append(m[L - 1][:, None] < m[-1][None, :])
    if nout > nin:
        masks[-1] = np.concatenate([masks[-1]] * int(nout / nin), axis=1)
    return [torch.from_numpy(mk.astype(np.uint8).T).float() for mk in masks]


def made_conditioner(x, sd, prefix, spec, context=None):
    """AutoregressiveConditioner.forward -> ConditionnalMADE.forward -> MADE.forward
    (AutoregressiveConditioner.py:24-25,108-109,135-141,150-151)."""
    cond_in = spec.get("cond_in", 0)
    inp = torch.cat((context, x), 1) if context is not None else x
    n_layers = len(spec["hidden"]) + 1
    y = inp
    p = f"{prefix}.masked_autoregressive_net.net"
    for l in range(n_layers):
        y = F.linear(y, sd[f"{p}.{2 * l}.mask"] * sd[f"{p}.{2 * l}.weight"], sd[f"{p}.{2 * l}.bias"])
        if l < n_layers - 1:
            y = torch.relu(y)
    out = y.view(inp.shape[0], -1, inp.shape[1]).permute(0, 2, 1)
    return out.contiguous()[:, cond_in:, :]


def coupling_conditioner(x, sd, prefix, spec, context=None):
    """CouplingConditioner.forward (CouplingConditioner.py:31-36)."""
    d = spec["d"]
    cond = int(d / 2)
    indep = d - cond
    if context is not None:
        x = torch.cat((x, context), 1)
    B = x.shape[0]
    h1 = sd[f"{prefix}.constants"].unsqueeze(0).expand(B, -1, -1)
    n_layers = len(spec["hidden"]) + 1
    h2 = mlp(x[:, :indep], sd, f"{prefix}.embeding_net.net", n_layers).view(B, cond, spec["out"])
    return torch.cat((h1, h2), 1)


# ----------------------------------------------------------------------------------------------
# Flow composition
# ----------------------------------------------------------------------------------------------
def flow_step(x, sd, k, spec, mode=None, noise=None, nb_steps=None, context=None):
    """NormalizingFlowStep.forward (NormalizingFlow.py:67-70) -> (z, logdet, h)."""
    pc = f"steps.{k}.conditioner"
    if spec["cond"] == "DAG":
        h = dag_conditioner(x, sd, pc, spec, mode, noise)
    elif spec["cond"] == "Autoregressive":
        h = made_conditioner(x, sd, pc, spec, context)
    elif spec["cond"] == "Coupling":
        h = coupling_conditioner(x, sd, pc, spec, context)
    else:
        raise ValueError(spec["cond"])
    if spec["norm"] == "affine":
        z, jac = affine_normalizer(x, h)
    else:
        S = spec["nb_steps"] if nb_steps is None else nb_steps
        z, jac = monotonic_normalizer(x, h, sd, f"steps.{k}.normalizer.integrand_net.net",
                                      len(spec["int_net"]) + 1, S)
    return z, torch.log(jac).sum(1), h


def flow_forward(x, sd, spec, modes=None, noises=None, nb_steps=None, context=None):
    """FCNormalizingFlow.forward (NormalizingFlow.py:118-126): columns reversed between steps,
    the last step's z returned un-reversed."""
    jac_tot = 0.
    inv_idx = torch.arange(x.shape[1] - 1, -1, -1).long()
    z = None
    for k in range(spec["nb_flow"]):
        z, jac, _ = flow_step(x, sd, k, spec, None if modes is None else modes[k],
                              None if noises is None else noises[k], nb_steps, context)
        x = z[:, inv_idx]
        jac_tot = jac_tot + jac
    return z, jac_tot


def constraints_loss(sd, spec, exponents=None):
    """FCNormalizingFlow.constraintsLoss (NormalizingFlow.py:128-132, :72-75)."""
    loss = 0.
    if spec["cond"] != "DAG":
        return loss
    for k in range(spec["nb_flow"]):
        p = spec["d"] % 50 if exponents is None else exponents[k]
        loss = loss + dag_loss(sd, f"steps.{k}.conditioner", p)
    return loss


def flow_loss(z, jac, sd, spec, exponents=None):
    """FCNormalizingFlow.loss (NormalizingFlow.py:144-146)."""
    return constraints_loss(sd, spec, exponents) - (jac + normal_log_density(z)).mean()


def compute_ll(x, sd, spec, modes=None, noises=None, nb_steps=None, context=None):
    """The ``compute_ll`` closure (ToyExperiments.py:134-137; UCIExperiments.py:159-160)."""
    z, jac = flow_forward(x, sd, spec, modes, noises, nb_steps, context)
    return normal_log_density(z) + jac, z


# ----------------------------------------------------------------------------------------------
# Reference-compatible parameter construction (for the CPU baseline / tests on the GPU box,
# where /root/reference does not exist)
# ----------------------------------------------------------------------------------------------
def _linear_init(out_f, in_f, gen):
    bound = 1. / math.sqrt(in_f)                     # == kaiming_uniform_(a=sqrt(5)) for nn.Linear
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def init_state_dict(spec, seed=0, A_prior=None):
    """Random-init weights with the reference constructors' *distributions* and key names
    (nn.Linear default init; A = 1.5 + 0.02 randn with zero diagonal, DAGConditioner.py:28,46-47;
    buffers DAGConditioner.py:49-61; constants ~ randn, CouplingConditioner.py:29)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    d = spec["d"]
    for k in range(spec["nb_flow"]):
        pc = f"steps.{k}.conditioner"
        if spec["cond"] == "DAG":
            A = A_prior.clone() if A_prior is not None else torch.ones(d, d) * 1.5 + torch.randn(d, d, generator=g) * .02
            A = A * (1. - torch.eye(d))
            sd[f"{pc}.A"] = A
            in_net = 2 * d if spec.get("hot_encoding", False) else d
            sizes = [in_net] + list(spec["hidden"]) + [spec["out"]]
            for l in range(len(sizes) - 1):
                w, b = _linear_init(sizes[l + 1], sizes[l], g)
                sd[f"{pc}.embedding_net.net.{2 * l}.weight"], sd[f"{pc}.embedding_net.net.{2 * l}.bias"] = w, b
            sd[f"{pc}.lambd"] = torch.tensor(0.)
            sd[f"{pc}.c"] = torch.tensor(1e-3)
            sd[f"{pc}.eta"] = torch.tensor(10.)
            sd[f"{pc}.gamma"] = torch.tensor(.9)
            sd[f"{pc}.l1_weight"] = torch.tensor(float(spec.get("l1", 0.)))
            sd[f"{pc}.dag_const"] = torch.tensor(1.)
            sd[f"{pc}.alpha"] = torch.tensor(1. / d)
            sd[f"{pc}.prev_trace"] = power_trace(A, 1. / d, d % 50)
        elif spec["cond"] == "Autoregressive":
            sizes = [d] + list(spec["hidden"]) + [spec["out"] * d]
            masks = made_masks(d, list(spec["hidden"]), spec["out"] * d)
            p = f"{pc}.masked_autoregressive_net.net"
            for l in range(len(sizes) - 1):
                w, b = _linear_init(sizes[l + 1], sizes[l], g)
                sd[f"{p}.{2 * l}.weight"], sd[f"{p}.{2 * l}.bias"], sd[f"{p}.{2 * l}.mask"] = w, b, masks[l]
        else:
            cond = int(d / 2)
            indep = d - cond
            sizes = [indep] + list(spec["hidden"]) + [spec["out"] * cond]
            for l in range(len(sizes) - 1):
                w, b = _linear_init(sizes[l + 1], sizes[l], g)
                sd[f"{pc}.embeding_net.net.{2 * l}.weight"], sd[f"{pc}.embeding_net.net.{2 * l}.bias"] = w, b
            sd[f"{pc}.constants"] = torch.randn(indep, spec["out"], generator=g)
        if spec["norm"] == "monotonic":
            sizes = [1 + spec["out"]] + list(spec["int_net"]) + [1]
            p = f"steps.{k}.normalizer.integrand_net.net"
            for l in range(len(sizes) - 1):
                w, b = _linear_init(sizes[l + 1], sizes[l], g)
                sd[f"{p}.{2 * l}.weight"], sd[f"{p}.{2 * l}.bias"] = w, b
    sd["z_log_density.pi"] = torch.tensor(math.pi)
    return sd


BUFFER_SUFFIXES = ("lambd", "c", "eta", "gamma", "l1_weight", "dag_const", "alpha", "prev_trace", "mask", "pi")


def trainable_keys(sd):
    return [k for k in sd if k.rsplit(".", 1)[-1] not in BUFFER_SUFFIXES]


def train_step_grads(x, sd, spec, modes=None, noises=None, nb_steps=None, exponents=None):
    """One forward + loss + backward on CPU; returns (loss, z, logdet, {key: grad})."""
    keys = trainable_keys(sd)
    sd2 = dict(sd)
    for k in keys:
        sd2[k] = sd[k].detach().clone().requires_grad_(True)
    z, jac = flow_forward(x, sd2, spec, modes, noises, nb_steps)
    loss = flow_loss(z, jac, sd2, spec, exponents)
    grads = torch.autograd.grad(loss, [sd2[k] for k in keys], allow_unused=True)
    return loss.detach(), z.detach(), jac.detach(), {k: g for k, g in zip(keys, grads)}
