"""GPU parity tests (-m gpu): the CUDA path, called through the product API -> autograd Functions -> C-ABI,
against (1) the reference-generated golden vectors, (2) the CPU oracle on seeded inputs at the BASELINE
configs' real layer shapes, (3) size-independent properties at BASELINE's full batch sizes.

Tolerances (BASELINE.json north_star, strict fp32 mode): per-sample log-likelihood 1e-4 relative,
per-parameter-tensor gradients 1e-3 relative L2."""
import ctypes as C

import pytest
import torch

import gnf_b200 as G
from helpers import golden_names
import parity

pytestmark = pytest.mark.gpu

LL_TOL, GRAD_TOL = 1e-4, 1e-3


def _mvo():
    import model_vs_oracle
    return model_vs_oracle


def test_product_library_is_the_device_build():
    assert G._lib.lib().gnf_has_device_code() == 1
    assert not G._lib._SIMULATOR


@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(name):
    parity.run_case(name, "cuda", LL_TOL, GRAD_TOL)


def _check(rep):
    bad = {k: v for k, v in rep.items() if (k.startswith("grad.") and not v < GRAD_TOL) or (k in ("ll", "loss") and not v < LL_TOL)}
    assert not bad, f"out of tolerance: {bad}\n{rep}"


@pytest.mark.parametrize("cfg,B", [("cfg1", 100), ("cfg2", 256), ("cfg3", 48), ("cfg4", 12), ("cfg5", 2)])
def test_train_step_vs_oracle_at_config_shapes(cfg, B):
    M = _mvo()
    _check(M.compare(M.CONFIGS[cfg], B, "cuda", train=True))


@pytest.mark.parametrize("cfg,B,S", [("cfg2", 300, 40), ("cfg3", 64, 40), ("cfg4", 20, 40), ("cfg4", 7, 150), ("cfg2", 33, 29)])
def test_compute_ll_vs_oracle_eval_steps(cfg, B, S):
    M = _mvo()
    _check(M.compare(M.CONFIGS[cfg], B, "cuda", train=False, nb_steps=S))


@pytest.mark.parametrize("mode", [dict(stoch_gate=False), dict(stoch_gate=False, s_thresh=False),
                                  dict(h_thresh=.3, stoch_gate=False), dict(h_thresh=.3)])
def test_dag_gate_modes_vs_oracle(mode):
    M = _mvo()
    _check(M.compare(M.CONFIGS["cfg2"], 64, "cuda", mode=mode, scaleA=.3, train=True))


def test_compute_ll_api_matches_forward():
    M = _mvo()
    model = M.build(M.CONFIGS["cfg2"], "cuda")
    parity.set_modes(model, dict(stoch_gate=False))
    x = torch.randn(128, 6, device="cuda")
    with torch.no_grad():
        ll, z = model.compute_ll(x)
        z2, jac = model(x)
        ll2 = model.z_log_density(z2) + jac
    assert torch.allclose(ll, ll2, rtol=1e-5, atol=1e-5) and torch.allclose(z, z2, rtol=1e-5, atol=1e-6)


# ---------------- kernel-level checks through the C-ABI wrappers ----------------
@pytest.mark.parametrize("M_,N,K", [(1, 1, 1), (127, 33, 65), (300, 630, 126), (1000, 30, 630), (64, 2, 1024), (513, 210, 21),
                                    # skinny layers (csrc/thin.cuh: one weight dimension <= 32, >= 512 rows): thin output / thin reduction
                                    (6300, 30, 630), (6300, 150, 30), (6301, 31, 1), (777, 32, 1536), (2049, 5, 33), (4096, 1, 700), (900, 700, 7),
                                    (6300, 630, 30), (513, 32, 32)])
def test_linear_engine_vs_torch(M_, N, K):
    g = torch.Generator(device="cuda").manual_seed(M_ + N + K)
    X = torch.randn(M_, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** .5
    b = torch.randn(N, device="cuda", generator=g)
    Y = G.ops.linear_fwd(X, W, b, relu=True)
    ref = torch.relu((X.double() @ W.double().t() + b.double())).float()
    assert torch.allclose(Y, ref, rtol=1e-4, atol=1e-5)
    dY = torch.randn(M_, N, device="cuda", generator=g)
    dX = G.ops.linear_dgrad(dY, N, W, X, M_)
    refdX = ((dY.double() @ W.double()) * (X > 0)).float()
    assert torch.allclose(dX, refdX, rtol=1e-4, atol=1e-4)
    dW = G.ops.linear_wgrad(dY, N, X, K, M_, N, K)
    refdW = (dY.double().t() @ X.double()).float()
    assert torch.allclose(dW, refdW, rtol=1e-4, atol=1e-3)
    db = G.ops.colsum(dY, N, M_, N).view(-1)
    assert torch.allclose(db, dY.double().sum(0).float(), rtol=1e-4, atol=1e-3)


def test_power_trace_golden_on_gpu():
    import os
    import numpy as np
    from helpers import GOLDEN, rel_l2
    f = np.load(os.path.join(GOLDEN, "power_trace.npz"))
    for key in sorted({k.rsplit(".", 1)[0] for k in f.files}):
        A = torch.from_numpy(f[key + ".A"]).cuda().requires_grad_(True)
        d, p, alpha = f[key + ".meta"]
        t = G.ops.PowerTraceFn.apply(A, float(alpha), int(p))
        t.backward()
        want = float(f[key + ".t"])
        assert abs(float(t.detach()) - want) <= 2e-5 * abs(want) + 1e-5 * float(d), key
        assert rel_l2(A.grad.cpu(), torch.from_numpy(f[key + ".dA"])) < 1e-4, key


def test_power_trace_large_d_vs_torch():
    """d = 784 (host-driven GEMM chain).  t = tr(B^p) - d cancels ~788 - 784 in fp32, so the forward value is
    compared on the trace's own scale (the reference computes the same fp32 difference)."""
    d, p = 784, 34
    A = (G.MNIST_A_prior(28, 2) * .7).cuda().requires_grad_(True)
    t = G.ops.PowerTraceFn.apply(A, 1. / d, p)
    t.backward()
    A2 = A.detach().double().requires_grad_(True)
    Bm = torch.eye(d, device="cuda", dtype=torch.float64) + A2 ** 2 / d
    t2 = torch.diag(torch.matrix_power(Bm, p)).sum() - d
    t2.backward()
    A3 = A.detach()
    t3 = torch.diag(torch.matrix_power(torch.eye(d, device="cuda") + A3 ** 2 / d, p)).sum() - d   # fp32 torch
    assert abs(float(t.detach()) - float(t2)) < 3e-5 * (d + abs(float(t2)))
    assert abs(float(t.detach()) - float(t3)) < 3e-5 * (d + abs(float(t2)))
    assert float((A.grad.double() - A2.grad).norm() / A2.grad.norm()) < 1e-4


def test_affine_clamps_h_in_place_and_masks_gradients():
    x = torch.randn(5, 3, device="cuda")
    h = (torch.randn(5, 3, 4, device="cuda") * 6).requires_grad_(True)
    hh = h * 1.
    z, jac = G.AffineNormalizer()(x, hh)
    (z.sum() + jac.sum()).backward()
    hc = h.detach()
    mu, ls = hc[..., 0].clamp(-5, 5), hc[..., 1].clamp(-5, 2)
    assert torch.equal(hh.detach()[..., 0], mu) and torch.equal(hh.detach()[..., 1], ls)       # in-place write-back (Q7)
    assert torch.allclose(z, x * ls.exp() + mu, rtol=1e-6, atol=1e-6)
    m0 = ((hc[..., 0] >= -5) & (hc[..., 0] <= 5)).float()
    m1 = ((hc[..., 1] >= -5) & (hc[..., 1] <= 2)).float()
    assert torch.allclose(h.grad[..., 0], m0) and torch.allclose(h.grad[..., 1], (x * ls.exp() + ls.exp()) * m1, rtol=1e-5, atol=1e-6)
    assert float(h.grad[..., 2:].abs().max()) == 0.


def test_wrappers_reject_cpu_and_wrong_dtype():
    with pytest.raises(TypeError):
        G.AffineNormalizer()(torch.randn(2, 3), torch.randn(2, 3, 2))
    with pytest.raises(TypeError):
        G.AffineNormalizer()(torch.randn(2, 3, device="cuda").double(), torch.randn(2, 3, 2, device="cuda").double())


def test_unsupported_shape_is_a_loud_error():
    with pytest.raises((RuntimeError, ValueError)):
        n = G.MonotonicNormalizer([300, 300], 4).cuda()          # hidden width > 256: outside the kernel's range
        n(torch.randn(4, 2, device="cuda"), torch.randn(4, 2, 4, device="cuda"))


# ---------------- properties at BASELINE's full sizes ----------------
@pytest.mark.parametrize("cfg,B", [("cfg2", 2500), ("cfg3", 10000), ("cfg4", 100), ("cfg1", 100)])
def test_full_size_batch_split_invariance_and_finiteness(cfg, B):
    """ll of a sample does not depend on which batch it is evaluated in (deterministic gate); everything finite."""
    M = _mvo()
    model = M.build(M.CONFIGS[cfg], "cuda")
    parity.set_modes(model, dict(stoch_gate=False))
    x = torch.randn(B, M.CONFIGS[cfg]["d"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    with torch.no_grad():
        ll, z = model.compute_ll(x)
        idx = torch.tensor([0, B // 3, B - 1], device="cuda")
        ll_s, z_s = model.compute_ll(x[idx].contiguous())
    assert torch.isfinite(ll).all() and torch.isfinite(z).all()
    assert torch.allclose(ll[idx], ll_s, rtol=2e-5, atol=1e-4)
    assert torch.allclose(z[idx], z_s, rtol=1e-5, atol=1e-5)


def test_full_size_quadrature_refinement_converges():
    """z(S) converges as S grows (CC quadrature of a smooth integrand): |z40 - z80| << |z20 - z80|, and the
    monotone map has a strictly positive Jacobian >= 0.05 (ELU+1.05)."""
    M = _mvo()
    model = M.build(M.CONFIGS["cfg4"], "cuda")
    parity.set_modes(model, dict(stoch_gate=False))
    x = torch.randn(100, 63, device="cuda")
    zs = {}
    with torch.no_grad():
        for S in (20, 40, 80):
            for n in model.getNormalizers():
                n.nb_steps = S
            zs[S], jac = model(x)
        h = model.getConditioners()[0](x)
        _, j = model.getNormalizers()[0](x, h)
    assert float(j.min()) >= 0.05 - 1e-6
    e20, e40 = float((zs[20] - zs[80]).abs().max()), float((zs[40] - zs[80]).abs().max())
    assert e40 <= e20 + 1e-6 and e40 < 1e-3


def test_one_step_affine_flow_inverts_exactly():
    M = _mvo()
    spec = dict(nb_flow=1, d=6, cond="DAG", hidden=[32, 32], out=2, hot_encoding=True, gumble_T=.5, l1=0., norm="affine")
    model = M.build(spec, "cuda")
    cond = model.getConditioners()[0]
    with torch.no_grad():
        cond.A.copy_(torch.tril(torch.ones(6, 6), -1).cuda())
    cond.post_process(.5)
    cond.is_invertible = True
    x = torch.randn(50, 6, device="cuda")
    with torch.no_grad():
        z, _ = model(x)
        xr = model.invert(z)
    assert float((x - xr).abs().max()) < 1e-4


def test_philox_gate_noise_is_reproducible_and_uniform():
    gs = G.ops.GateSpec(G._lib.GATE_GUMBEL, G._lib.IMP_SOFT, 0., .5, seed=1234, offset=7)
    a1, a2 = G.ops.dag_dump_noise(gs, 64, 20, "cuda")
    b1, b2 = G.ops.dag_dump_noise(gs, 64, 20, "cuda")
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    assert 0. < float(a1.min()) and float(a1.max()) < 1.
    assert abs(float(a1.mean()) - .5) < .02 and abs(float(a2.mean()) - .5) < .02
    assert abs(float(((a1 - .5) * (a2 - .5)).mean())) < .01


def test_philox_uniforms_stay_strictly_inside_unit_interval():
    """33M draws: the generator must never return 0 or 1 (either makes the Gumbel gate infinite -> NaN loss)."""
    gs = G.ops.GateSpec(G._lib.GATE_GUMBEL, G._lib.IMP_SOFT, 0., .5, seed=99, offset=3)
    a1, a2 = G.ops.dag_dump_noise(gs, 4096, 64, "cuda")
    for a in (a1, a2):
        assert float(a.min()) > 0. and float(a.max()) < 1.
        g = -torch.log(-torch.log(a))
        assert torch.isfinite(g).all()


@pytest.mark.parametrize("cfg,B,steps", [("cfg4", 100, 60), ("cfg2", 2500, 40), ("cfg5", 16, 6)])
def test_training_stays_finite(cfg, B, steps):
    """A short Adam run at the reference's hyper-parameters with the in-kernel Philox gate: loss and parameters stay finite."""
    M = _mvo()
    model = M.build(M.CONFIGS[cfg], "cuda")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    g = torch.Generator(device="cuda").manual_seed(0)
    for it in range(steps):
        x = torch.randn(B, M.CONFIGS[cfg]["d"], device="cuda", generator=g)
        opt.zero_grad()
        z, jac = model(x)
        loss = model.loss(z, jac)
        loss.backward()
        opt.step()
        assert torch.isfinite(loss.detach()), f"loss became {float(loss.detach())} at step {it}"
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_cuda_graph_training_step_matches_eager():
    """GraphedTrainStep replays the same step the eager path runs: same parameters after the same batches
    (deterministic gate; only atomics' summation order differs)."""
    M = _mvo()
    spec = M.CONFIGS["cfg2"]
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(256, 6, device="cuda", generator=g) for _ in range(6)]

    def make():
        model = M.build(spec, "cuda", seed=3)
        parity.set_modes(model, dict(stoch_gate=False))
        bucket = G.dist.GradBucket(model.parameters())
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, fused=True, capturable=True)
        return model, bucket, opt

    model_e, bucket_e, opt_e = make()
    for x in [xs[0]] * 3 + xs[1:]:
        bucket_e.zero()
        z, jac = model_e(x)
        model_e.loss(z, jac).backward()
        opt_e.step()
    model_g, bucket_g, opt_g = make()
    step = G.GraphedTrainStep(model_g, opt_g, bucket_g, xs[0], allreduce=False, warmup=3)
    losses = [float(step(x)) for x in xs[1:]]
    assert all(l == l for l in losses)
    for (k, pe), (_, pg) in zip(model_e.named_parameters(), model_g.named_parameters()):
        err = float((pe - pg).norm() / pe.norm().clamp_min(1e-12))
        assert err < 2e-3, (k, err)     # Adam amplifies the atomics-order noise of 8 steps; a wrong graph would be O(1) off


def test_cuda_graph_replays_draw_fresh_gate_noise():
    M = _mvo()
    model = M.build(M.CONFIGS["cfg2"], "cuda", seed=3)
    bucket = G.dist.GradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=0., fused=True, capturable=True)
    x = torch.randn(64, 6, device="cuda")
    step = G.GraphedTrainStep(model, opt, bucket, x, allreduce=False, warmup=3)
    l1, l2 = float(step(x)), float(step(x))
    assert l1 != l2, "two replays on the same batch with lr=0 must differ through the stochastic gate noise"


@pytest.mark.parametrize("cfg,B", [("cfg4", 100), ("cfg2", 300), ("cfg1", 100)])
def test_cuda_graph_eval_step_matches_eager_and_follows_weight_updates(cfg, B):
    """GraphedEvalStep replays compute_ll: same per-sample ll as the eager call on new batches, and -- the padded / split
    weight copies are made inside the graph -- also after the parameters were updated in place."""
    M = _mvo()
    spec = M.CONFIGS[cfg]
    model = M.build(spec, "cuda", seed=5)
    parity.set_modes(model, dict(stoch_gate=False))
    g = torch.Generator(device="cuda").manual_seed(2)
    xs = [torch.randn(B, spec["d"], device="cuda", generator=g) for _ in range(3)]
    step = G.GraphedEvalStep(model, xs[0])
    for it, x in enumerate(xs):
        if it == 2:
            with torch.no_grad():
                for p in model.parameters():
                    p.mul_(1.01)
        ll_g, z_g = step(x)
        ll_g, z_g = ll_g.clone(), z_g.clone()
        with torch.no_grad():
            ll_e, z_e = model.compute_ll(x)
        assert float(((ll_g - ll_e).abs() / ll_e.abs().clamp_min(1e-6)).max()) < 1e-5, it
        assert float((z_g - z_e).abs().max()) < 1e-4, it


def test_umnn_backward_saved_activations_equals_recompute():
    """The backward that reloads the forward's hidden activations and the one that recomputes them (UMNN's way) give the
    same gradients (bit-identical activations; only the atomics' order differs)."""
    M = _mvo()
    model = M.build(M.CONFIGS["cfg4"], "cuda", seed=2)
    parity.set_modes(model, dict(stoch_gate=False))
    x = torch.randn(40, 63, device="cuda")
    grads = []
    old = G.ops.SAVE_ACTIVATIONS_MAX_BYTES
    try:
        for budget in (16 << 30, 0):
            G.ops.SAVE_ACTIVATIONS_MAX_BYTES = budget
            model.zero_grad()
            z, jac = model(x)
            model.loss(z, jac).backward()
            grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
    finally:
        G.ops.SAVE_ACTIVATIONS_MAX_BYTES = old
    for k in grads[0]:
        err = float((grads[0][k] - grads[1][k]).norm() / grads[1][k].norm().clamp_min(1e-20))
        assert err < 1e-5, (k, err)


def test_dag_loss_fused_matches_reference_formula():
    """DAGConditioner.loss (DAGConditioner.py:268-271): fused kernel pair vs the reference's torch expression, values and
    gradients, including a dual-variable state in the middle of training."""
    torch.manual_seed(3)
    d = 17
    A = (torch.randn(d, d, device="cuda") * .7).requires_grad_(True)
    t = torch.tensor(3.25, device="cuda", requires_grad=True)
    for lambd, c, dag_const, l1 in ((0., 1e-3, 1., .1), (2.5, 10., 1., 0.), (1., 1., 0., .3)):
        duals = [torch.tensor(v, device="cuda") for v in (lambd, c, dag_const, l1)]
        out = G.ops.DagLossFn.apply(A, t, *duals)
        gA, gt = torch.autograd.grad(out * 1.7, (A, t))
        ref = duals[2] * (duals[0] * t + duals[1] / 2 * t ** 2) + duals[3] * A.abs().mean()
        rA, rt = torch.autograd.grad(ref * 1.7, (A, t))
        assert abs(float(out) - float(ref)) <= 1e-6 * max(1., abs(float(ref)))
        assert float((gA - rA).abs().max()) <= 1e-7 and abs(float(gt - rt)) <= 1e-6 * max(1., abs(float(rt)))
    # fp32 overflow of t^2 must survive
    big = torch.tensor(3e20, device="cuda")
    out = G.ops.DagLossFn.apply(A.detach(), big, *[torch.tensor(v, device="cuda") for v in (0., 1e-3, 1., 0.)])
    assert torch.isinf(out)


# ---------------------------------------------------------------------------------------------------------------
# Edge batches: single sample, ragged sizes around the 64 / 128-row tile boundaries, and the empty batch, through every
# engine combination of the training path (fused FFMA, layer-wise with FFMA GEMMs, layer-wise with tensor-core GEMMs).
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def engines():
    def set_(umnn, gemm):
        G.ops.UMNN_ENGINE = umnn
        G.ops.set_gemm_mode(gemm)
    yield set_
    G.ops.UMNN_ENGINE = "auto"
    G.ops.set_gemm_mode("ffma")


@pytest.mark.parametrize("umnn,gemm", [("fused", "ffma"), ("layerwise", "ffma"), ("layerwise", "tf32x3"), ("auto", "auto")])
@pytest.mark.parametrize("cfg,B", [("cfg2", 1), ("cfg2", 22), ("cfg4", 1), ("cfg4", 3), ("cfg3", 7)])
def test_edge_batches_vs_oracle(engines, umnn, gemm, cfg, B):
    M = _mvo()
    engines(umnn, gemm)
    _check(M.compare(M.CONFIGS[cfg], B, "cuda", train=True))


@pytest.mark.parametrize("umnn,gemm", [("fused", "ffma"), ("layerwise", "tf32x3")])
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_empty_batch(engines, umnn, gemm, cfg):
    M = _mvo()
    engines(umnn, gemm)
    spec = M.CONFIGS[cfg]
    model = M.build(spec, "cuda")
    x = torch.empty(0, spec["d"], device="cuda")
    z, jac = model(x)
    assert z.shape == (0, spec["d"]) and jac.shape == (0,)
    with torch.no_grad():
        ll, _ = model.compute_ll(x)
    assert ll.shape == (0,)
    torch.cuda.synchronize()
