#!/usr/bin/env python
"""Benchmark of the Graphical-Normalizing-Flows hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg1..cfg5] [--batch B_per_gpu] [--mode train|eval]

One "step" = one pass of the hot path over one batch of synthetic input:
  train: zero_grad -> forward -> loss -> backward -> (gradient all-reduce if N>1) -> Adam step
  eval : compute_ll under no_grad.
Default workload: BASELINE.json configs[3] (BSDS300 shape, d=63, DAG conditioner + Monotonic/UMNN normalizer),
the configuration north_star's target is quoted on; 100 samples per GPU (its YAML batch), weak-scaled.

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM, timed per step with CUDA
events (L2 flushed between timed steps), max over ranks.  `e2e` = the same metric through the public API with
pinned HOST input batches copied H2D and the loss read back D2H inside the timed region.
`--impl reference` times the reference ITSELF on the host CPU (all threads) on the same workload: the verbatim copy of its
`models/` package under the git-ignored baseline/_ref/ (made by baseline/install_ref.py in the build container, shipped
with the snapshot), stock classes and code path, with oracle/UMNN.py standing in for the uninstallable UMNN==1.0
("kind": "reference+restated-UMNN"); only if that copy is missing does it fall back to the oracle port ("kind": "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gnf_b200 as G  # noqa: E402

CONFIGS = G.CONFIGS

DEFAULT_BATCH = {"cfg1": 100, "cfg2": 2500, "cfg3": 10000, "cfg4": 100, "cfg5": 100}
WORKLOAD_NAME = {
    "cfg1": "toy-8gaussians d=2, 3x(DAG+Affine)",
    "cfg2": "UCI-POWER shape d=6, DAG+Monotonic(UMNN) S=20",
    "cfg3": "UCI-HEPMASS shape d=21, Autoregressive+Monotonic(UMNN) S=20",
    "cfg4": "BSDS300 shape d=63, DAG(630x3,hot)+Monotonic(UMNN 150x3) S=20",
    "cfg5": "MNIST shape d=784, DAG(1024x3,hot,prior k=2)+Affine",
}
ADAM = {"cfg1": (1e-4, 1e-5), "cfg2": (1e-3, 1e-5), "cfg3": (1e-3, 1e-4), "cfg4": (1e-3, 1e-4), "cfg5": (1e-3, 1e-5)}


# ------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------
def flops_per_sample(spec, S):
    d = spec["d"]
    hid = list(spec["hidden"])
    H = spec["out"]
    if spec["cond"] == "DAG":
        sizes = [d] + hid + [H]                      # one-hot half of layer 1 is a bias gather: not counted
        f_cond = 2 * d * sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))
    elif spec["cond"] == "Autoregressive":
        sizes = [d] + hid + [H * d]
        f_cond = 2 * sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))
    else:
        c = d // 2
        sizes = [d - c] + hid + [H * c]
        f_cond = 2 * sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))
    f_int = 0
    if spec["norm"] == "monotonic":
        I = list(spec["int_net"])
        per_node = I[0] + sum(a * b for a, b in zip(I[:-1], I[1:])) + I[-1]
        f_int = 2 * d * ((S + 2) * per_node + H * I[0])
    f_cond *= spec["nb_flow"]
    f_int *= spec["nb_flow"]
    return f_cond, f_int


def umnn_kernel_flops(spec, S, rows, backward):
    """FLOPs one gnf_umnn_{fwd,bwd} call executes algorithmically for `rows` = B*d rows."""
    I = list(spec["int_net"])
    E = spec["out"]
    per_node = (1 + E) * I[0] + sum(a * b for a, b in zip(I[:-1], I[1:])) + I[-1]
    if not backward:
        return 2 * rows * (S + 1) * per_node
    return 3 * 2 * rows * (S + 2) * per_node          # recompute forward + dgrad + wgrad


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region: NVML in-process every 5 ms (nvidia-smi, one
    process spawn per sample, gets a single sample into a 50 ms region), nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _reasons_nvml(self):
        n = self.nvml
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        return [("Active" if mask & bits[k] else "Not Active") for k in self.NAMES]

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    sm = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    self.samples.append([str(sm), str(self.max_sm)] + self._reasons_nvml())
                    time.sleep(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(self.NAMES, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref/models, stock classes and code path, + the restated UMNN dependency) when the
# copy made by baseline/install_ref.py travelled with the snapshot; else the oracle port (reference algorithm, torch CPU)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(cfg, B, mode, S):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_arm
    lr, wd = ADAM[cfg]
    if ref_arm.available():
        return ref_arm.step_fn(CONFIGS[cfg], B, mode, S, lr, wd), "reference+restated-UMNN"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gnf_oracle as O
    spec = {k: v for k, v in CONFIGS[cfg].items() if k != "A_prior"}
    A_prior = None
    if CONFIGS[cfg].get("A_prior") == "mnist":
        A_prior = G.MNIST_A_prior(28, 2)
    sd = O.init_state_dict(spec, seed=0, A_prior=A_prior)
    keys = O.trainable_keys(sd)
    for k in keys:
        sd[k] = sd[k].clone().requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in keys], lr=lr, weight_decay=wd)
    g = torch.Generator().manual_seed(0)
    d = spec["d"]

    def noises():
        if spec["cond"] != "DAG":
            return None
        return [(torch.rand(B, d, d), torch.rand(B, d, d)) for _ in range(spec["nb_flow"])]

    def train_step():
        x = torch.randn(B, d, generator=g)
        opt.zero_grad()
        z, jac = O.flow_forward(x, sd, spec, None, noises(), S)
        loss = O.flow_loss(z, jac, sd, spec)
        loss.backward()
        opt.step()
        return float(loss.detach())

    def eval_step():
        x = torch.randn(B, d, generator=g)
        with torch.no_grad():
            ll, _ = O.compute_ll(x, sd, spec, None, noises(), S)
        return float(ll.mean())

    return (train_step if mode == "train" else eval_step), "port"


def time_cpu(cfg, B, mode, S, steps, warmup, budget_s=25.):
    """Bounded CPU timing: EXACTLY `steps` timed steps after `warmup` untimed ones; the batch is reduced (stated in `sample`)
    when full-batch steps would blow the time budget."""
    torch.set_num_threads(os.cpu_count() or 1)
    Bc = B
    probe_B = min(B, 16)
    fn, kind = cpu_reference_step_fn(cfg, probe_B, mode, S)
    fn()
    t0 = time.perf_counter(); fn(); t1 = time.perf_counter()
    per_sample = (t1 - t0) / probe_B
    total_steps = steps + warmup
    if per_sample * B * total_steps > budget_s:
        Bc = max(1, int(budget_s / (per_sample * total_steps)))
    fn, kind = cpu_reference_step_fn(cfg, Bc, mode, S)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    mean = sum(ts) / len(ts)
    ts.sort()
    med = ts[len(ts) // 2]
    what = ("the reference's own classes (baseline/_ref/models: buildFCNormalizingFlow, forward, loss, backward, Adam) with "
            "oracle/UMNN.py standing in for the absent UMNN==1.0" if kind != "port" else "the oracle port of the reference algorithm")
    return {"value": Bc / mean, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "steps": steps, "warmup": warmup, "batch": Bc, "median_step_ms": med * 1e3, "mean_step_ms": mean * 1e3,
            "sample": f"{steps} {mode} steps of {WORKLOAD_NAME[cfg]} at batch {Bc} (of {B}) on the host CPU after {warmup} warm-up "
                      f"steps, {what}; value = batch / mean step time"}, mean


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the config's own batch)")
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--nb-steps", type=int, default=None, help="quadrature steps S (default 20 train / 40 eval)")
    ap.add_argument("--precision", default="strict", choices=["strict", "tf32"],
                    help="strict: fp32 FFMA kernels; tf32: tensor-core (tcgen05) UMNN forward, ll tolerance 2e-3")
    ap.add_argument("--gemm", default="auto", choices=["ffma", "tf32x3", "tf32", "auto", "auto-fast"],
                    help="conditioner GEMM engine: fp32 FFMA, tensor-core 3xTF32 (fp32-equivalent) or single-pass TF32")
    ap.add_argument("--umnn-engine", default="auto", choices=["auto", "fused", "layerwise"],
                    help="strict UMNN integral: fused FFMA kernels, layer-wise passes on the GEMM engine, or auto")
    ap.add_argument("--random-steps", action="store_true",
                    help="train: S = nb_steps + U{0..9} per step as the reference's UCI driver does (UCIExperiments.py:131-133); one "
                         "captured graph per S")
    ap.add_argument("--no-side-branch", action="store_true", help="measurement: keep the DAG penalty chain in line with the forward inside the captured step")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "overlap", "single", "none"],
                    help="N > 1, gradient average of the flat bucket after the backward: peer = ONE libgnf kernel over NVLink peer memory (default; falls back "
                         "to single when CUDA IPC is unavailable); single = one NCCL all-reduce; overlap = NCCL sub-buckets overlapped with the rest of "
                         "the backward (measured slower: the NCCL CTAs displace CTAs of the 148-CTA persistent GEMMs); none = measurement only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true", help="train mode: skip the additional log-lik eval measurement")
    ap.add_argument("--cuda-graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step (incl. the NCCL gradient all-reduce) as one captured CUDA graph (auto = on)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = args.config
    spec = CONFIGS[cfg]
    B = args.batch or DEFAULT_BATCH[cfg]
    S = args.nb_steps or (20 if args.mode == "train" else 40)
    metric = "train_samples_per_s" if args.mode == "train" else "loglik_eval_samples_per_s"
    config = {"workload": WORKLOAD_NAME[cfg], "config": cfg, "batch_per_gpu": B, "global_batch": B * max(world, 1),
              "nb_steps": S, "mode": args.mode, "parallelism": f"dp{max(world, 1)}", "gate": "stochastic (reference default)",
              "precision_mode": ("strict fp32-equivalent (FFMA kernels + 3xTF32 tcgen05 GEMMs; ll 1e-4 / gradients 1e-3 vs the oracle)"
                                 if args.precision == "strict" else
                                 "tf32 tensor-core UMNN forward (tcgen05, ll tol 2e-3) + strict fp32 elsewhere"), "l2": "flushed between timed steps (256 MiB write)"}

    config["gemm_engine"] = args.gemm
    config["umnn_engine"] = args.umnn_engine
    config["cuda_graph"] = args.cuda_graph in ("on", "auto")
    config["allreduce"] = args.allreduce
    if args.random_steps:
        config["nb_steps"] = f"{S}+U{{0..9}} per step (UCIExperiments.py:131-133)"

    # ---------------- reference arm ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        cb, mean = time_cpu(cfg, B, args.mode, S, args.steps, args.warmup, budget_s=150.)
        line = {"metric": metric, "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = G.build_from_spec(spec, dev, seed=0)
    G.dist.broadcast_parameters(model)
    G.dist.decorrelate_gate_noise(model, rank)
    d = spec["d"]
    lr, wd = ADAM[cfg]
    bucket = G.dist.GradBucket(model.parameters(), overlap=(args.allreduce == "overlap"), peer=(args.allreduce == "peer"))
    if args.allreduce == "peer":
        config["allreduce"] = "peer" if bucket.peer is not None else ("single" if world > 1 else "peer")
    # the reference's optimizer (torch.optim.Adam(lr, weight_decay), UCIExperiments.py:100) as one multi-tensor launch per step
    opt = G.FusedAdam(model.parameters(), lr=lr, weight_decay=wd)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_pool = 8
    pool = [torch.randn(B, d, device=dev, generator=gen) for _ in range(n_pool)]
    host_pool = [torch.randn(B, d).pin_memory() for _ in range(n_pool)]
    flush_buf = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the roofline kernels, from the committed
    # `ncu --set full` captures: profiles/ncu_traffic.json {"<kernel>@<cfg>": bytes}
    traffic_table = {}
    try:
        traffic_table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass

    def train_step(x):
        bucket.begin_step()
        z, jac = model(x)
        loss = model.loss(z, jac)
        loss.backward()
        if args.allreduce != "none":      # "none" = measurement only: ranks train independently (how much of the N-GPU step is the collective)
            bucket.finish_step()
        opt.step()
        return loss

    def eval_step(x):
        with torch.no_grad():
            ll, _ = model.compute_ll(x)
        return ll.mean()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(mode, precision, gemm, S_, steps, warmup, sample_clocks, use_graph=False):
        """One arm: W warm-up steps, K device-timed steps (CUDA events per step, L2 flushed between steps), then K
        end-to-end steps (pinned host batch -> H2D -> step -> loss D2H).  Max over ranks.
        use_graph: the timed steps replay ONE captured CUDA graph of the whole training step; the per-kernel times for
        the roofline object are then taken from a short eager pass first (a replay has no per-launch host hooks)."""
        G.ops.set_gemm_mode(gemm)
        G.ops.UMNN_ENGINE = args.umnn_engine
        for n in model.getNormalizers():
            if hasattr(n, "nb_steps"):
                n.nb_steps = S_
                n.precision = precision
        step = train_step if mode == "train" else eval_step
        for i in range(warmup):
            step(pool[i % n_pool])
        barrier()
        eager_ktimes, eager_kflops, eager_kshapes, eager_launches, n_eager = None, None, None, 0, 0
        if use_graph:
            l0 = G.ops.launch_count()
            G.ops.enable_kernel_timing(True)
            n_eager = min(steps, 10)
            for i in range(n_eager):
                # keep the launch queue full: a spin kernel (~4 ms) ahead of the step lets the host enqueue every launch of the step
                # before the first one starts, so that an event pair brackets its kernel(s) and not the host's launch gaps
                # (idle-queue pairs read 10-14 us for a 3-us kernel)
                torch.cuda._sleep(8_000_000)
                step(pool[i % n_pool])
            eager_ktimes = G.ops.collect_kernel_timing()
            eager_kflops = G.ops.collect_call_flops()
            eager_kshapes = G.ops.collect_call_shapes()
            G.ops.enable_kernel_timing(False)
            eager_launches = (G.ops.launch_count() - l0) // n_eager
            if mode == "train" and args.random_steps:
                # the reference draws S per batch: the graph is keyed by S (GraphedTrainStep recaptures when nb_steps changes; here
                # ten graphs are kept instead of recapturing)
                import random
                rnd = random.Random(1234 + rank)
                graphs = {}

                def step(x):
                    Sx = S_ + rnd.randrange(10)
                    for n in model.getNormalizers():
                        if hasattr(n, "nb_steps"):
                            n.nb_steps = Sx
                    if Sx not in graphs:
                        graphs[Sx] = G.GraphedTrainStep(model, opt, bucket, pool[0], allreduce=args.allreduce != "none", warmup=2, side_branch=not args.no_side_branch)
                    return graphs[Sx](x)
                for Sx in range(S_, S_ + 10):             # capture outside the timed region
                    for n in model.getNormalizers():
                        if hasattr(n, "nb_steps"):
                            n.nb_steps = Sx
                    graphs[Sx] = G.GraphedTrainStep(model, opt, bucket, pool[0], allreduce=args.allreduce != "none", warmup=2, side_branch=not args.no_side_branch)
            elif mode == "train":
                step = G.GraphedTrainStep(model, opt, bucket, pool[0], allreduce=args.allreduce != "none", warmup=3, side_branch=not args.no_side_branch)
            else:
                graphed_eval = G.GraphedEvalStep(model, pool[0], warmup=3)
                step = lambda x: graphed_eval(x)[0].mean()
            for i in range(3):
                step(pool[i % n_pool])
            barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
        l0 = G.ops.launch_count()
        if not use_graph:
            G.ops.enable_kernel_timing(True)
        evs = []
        barrier()
        for i in range(steps):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(pool[i % n_pool])
            e1.record()
            evs.append((e0, e1))
        barrier()
        if use_graph:
            launches, ktimes, kflops, kshapes = eager_launches * steps, eager_ktimes, eager_kflops, eager_kshapes
        else:
            launches = G.ops.launch_count() - l0
            ktimes = G.ops.collect_kernel_timing()
            kflops = G.ops.collect_call_flops()
            kshapes = G.ops.collect_call_shapes()
            G.ops.enable_kernel_timing(False)
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        barrier()
        t0 = time.perf_counter()
        last = 0.
        for i in range(steps):
            x = host_pool[i % n_pool].to(dev, non_blocking=True)
            last = float(step(x).detach())
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        t = torch.tensor([dev_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        f_cond, f_int = flops_per_sample(spec, S_)
        flops_step = (3 * f_cond + 4 * f_int if mode == "train" else f_cond + f_int) * B
        # ---- roofline of the dominant kernel: the candidate with the largest total time per step.
        #   tc_gemm_kernel   every gnf_linear_{fwd,dgrad,wgrad}_tc call is ONE launch of it (+ a memset for wgrad); FLOPs per
        #                    call are noted by the op wrappers (2*M*N*K)
        #   UMNN entries     one fused kernel (gnf_umnn_fwd / _bwd / _fwd_tc) or the layer-wise composite (_lw: per-layer
        #                    launches inside one C-ABI call; its hidden GEMMs run on rw_gemm_kernel / tc_gemm_kernel)
        n_timed = n_eager if use_graph else steps
        step_ms = dev_ms / steps
        cands = []
        tc_names = [k for k in ("gnf_linear_fwd_tc", "gnf_linear_dgrad_tc", "gnf_linear_wgrad_tc") if ktimes.get(k)]
        single = precision == "tf32" or gemm in ("tf32", "auto-fast")
        by_kernel = {}                                 # the engine's kernel that each timed C-ABI call launched -> [(ms, flops)]
        for k in tc_names:
            shapes, fls = kshapes.get(k, []), kflops.get(k, [])
            if len(shapes) != len(ktimes[k]) or len(fls) != len(ktimes[k]):
                continue
            for ms_, fl_, shp in zip(ktimes[k], fls, shapes):
                kern = G.ops.tc_kernel_name(k.split("_")[2], *shp, passes=(1 if single else 3)) if shp else "tc_gemm_kernel"
                by_kernel.setdefault(kern, []).append((ms_, fl_, k))
        tc_notes = {
            "tc_gemm2_kernel": "tcgen05 kind::tf32 GEMM engine v2 (conditioner layers, forward + dgrad): 3xTF32 = fp32-equivalent (3 tensor-core "
                               "passes per algorithmic FLOP: the tensor pipe sees 3x the achieved figure, against a TF32 dense peak of half the "
                               "bf16 one); activation operand split in registers and fed through TMEM, weights pre-split once per call",
            "tc_wgrad2_kernel": "tcgen05 kind::tf32 GEMM engine v2, weight gradient (split-K over the batch rows, 3xTF32): dY^T through registers "
                                "into TMEM, X split in shared memory; the call also zero-fills dW",
            "tc_gemm_kernel": "tcgen05 kind::tf32 GEMM engine (first engine: both operands staged and split in shared memory), "
                              + ("single-pass TF32" if single else "3xTF32 = fp32-equivalent: 3 tensor-core passes per algorithmic FLOP")}
        for kern, rows in by_kernel.items():
            ms, fl, n = sum(r[0] for r in rows), sum(r[1] for r in rows), len(rows)
            if fl > 0:
                cands.append(dict(kernel=kern, entries=sorted({r[2] for r in rows}), flops_per_launch=fl / n, avg_launch_ms=ms / n,
                                  launches_per_step=n / n_timed, ms_per_step=ms / n_timed, note=tc_notes[kern]))
        if spec["cond"] == "DAG":
            # DAG layer 1 (K1): gnf_dag_l1_{fwd,wgrad,dgrad} = one kernel each, 2*B*d*d*H1 FLOPs (the one-hot half is a bias gather)
            H1 = spec["hidden"][0]
            for kname, kern in (("gnf_dag_l1_fwd", "dag_l1_fwd_kernel"), ("gnf_dag_l1_wgrad", "dag_l1_wgrad_kernel"), ("gnf_dag_l1_dgrad", "dag_l1_dgrad_kernel")):
                if not ktimes.get(kname):
                    continue
                ms, n = sum(ktimes[kname]), len(ktimes[kname])
                per_step = n / n_timed              # 1 = the resident-gate FFMA kernel; 2 = tensor-core GEMM against the saved gate planes + reduction
                if per_step > 1.5:
                    kern = "tc_gemm_kernel+dag_l1_reduce_narrow_kernel" if kname.endswith("dgrad") else kern
                elif kname != "gnf_dag_l1_fwd" and d <= 64 and gemm in ("auto", "tf32x3") and mode == "train":
                    kern = "tc_gemm_kernel"         # layer-1 weight gradient on the tensor-core engine against the saved plane (one launch)
                cands.append(dict(kernel=kern, entries=[kname], flops_per_launch=2. * B * d * d * H1 / max(per_step, 1.), avg_launch_ms=ms / n,
                                  launches_per_step=per_step, ms_per_step=ms / n_timed,
                                  note="DAG masked embedding fused into layer 1 (forward: strict-fp32 FFMA kernel, the gate -- Philox + Gumbel "
                                       "sigmoid -- generated on chip; narrow-flow backward: 3xTF32 tensor-core GEMMs against the gate planes the "
                                       "forward kept); the fraction is quoted against the measured bf16 tensor peak as the contract asks"))
        if spec["norm"] == "monotonic":
            for kname, bwd in (("gnf_umnn_bwd", True), ("gnf_umnn_bwd_lw", True), ("gnf_umnn_fwd", False), ("gnf_umnn_fwd_lw", False),
                               ("gnf_umnn_fwd_tc", False), ("gnf_umnn_fwd_tc3", False)):
                if not ktimes.get(kname) or (mode == "train" and not bwd and kname != "gnf_umnn_fwd_tc3"):
                    continue
                ms = sum(ktimes[kname])
                n = len(ktimes[kname])
                if kname == "gnf_umnn_fwd_tc3":
                    nodes = S_ + (2 if mode == "train" else 1)
                    I = list(spec["int_net"])
                    fl = 2. * B * d * (nodes * (I[0] + sum(a * b for a, b in zip(I[:-1], I[1:])) + I[-1]) + spec["out"] * I[0])
                    cands.append(dict(kernel="umnn_fwd_tc3_kernel", entries=[kname], flops_per_launch=fl, avg_launch_ms=ms / n,
                                      launches_per_step=n / n_timed, ms_per_step=ms / n_timed,
                                      note="fused strict UMNN forward: tcgen05 kind::tf32 in 3xTF32 (3 tensor-core passes per algorithmic "
                                           "FLOP), activation chain resident in TMEM, hidden weights streamed as pre-split hi/lo K-chunks; "
                                           "the call also runs the pack kernel and the once-per-row conditioning GEMM; training stores "
                                           "the three activation planes (267 MB at cfg4 B=100) for the layer-wise backward"))
                    continue
                note = ("tcgen05 kind::tf32 kernel, activation chain resident in TMEM (the TF32 dense peak is half the bf16 peak the "
                        "fraction is quoted against)" if kname.endswith("_tc") else
                        "layer-wise engine: ONE C-ABI call = per-layer launches (hidden GEMMs on tcgen05 in 3xTF32: rw_gemm_kernel for "
                        "forward/dgrad, tc_gemm_kernel for wgrad; elementwise/reduction passes between them); duration is the whole call"
                        if kname.endswith("_lw") else
                        "strict-fp32 FFMA kernel (no tensor-core instructions): fp32 CUDA-core ceiling is ~74 TFLOP/s; fraction is "
                        "quoted against the measured bf16 tensor peak as the contract asks")
                cands.append(dict(kernel=kname, entries=[kname], flops_per_launch=umnn_kernel_flops(spec, S_, B * d, backward=bwd),
                                  avg_launch_ms=ms / n, launches_per_step=n / n_timed, ms_per_step=ms / n_timed, note=note))
        roofline, others = None, []
        # a layer-wise composite call is several kernels: it is reported next to the kernels, never as "the" dominant kernel
        # when a single-kernel candidate exists (its hidden GEMMs are tc_gemm_kernel / rw_gemm_kernel launches themselves)
        for c in sorted(cands, key=lambda c: (c["kernel"].endswith("_lw") and any(not o["kernel"].endswith("_lw") for o in cands),
                                              -c["ms_per_step"])):
            ach = c["flops_per_launch"] / (c["avg_launch_ms"] / 1e3) / 1e12
            obj = {"bound": "tensor", "kernel": c["kernel"], "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                   "traffic": traffic_table.get(f"{c['kernel']}.{mode}@{cfg}", traffic_table.get(f"{c['kernel']}@{cfg}")),
                   "avg_launch_ms": c["avg_launch_ms"],
                   "flops_per_launch": c["flops_per_launch"], "launches_per_step": c["launches_per_step"], "peak_source": peak_src,
                   "note": c["note"], "share_of_step": c["ms_per_step"] / step_ms, "timed_entry_points": c["entries"]}
            if roofline is None:
                roofline = obj
            else:
                others.append({k: obj[k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "launches_per_step", "share_of_step")})
        if roofline is not None and others:
            roofline["other_kernels"] = others
        # ---- K4 (HBM-bound elementwise + row reductions): algorithmic bytes / CUDA-event duration of the call (forward
        #      16*B*d + 8*B bytes, backward the same order) against the measured copy bandwidth
        peak_gbs = peaks.get("hbm_gbs", 6650.)
        k4 = []
        for kname, nbytes in (("gnf_affine_fwd", 16. * B * d + 8. * B), ("gnf_affine_bwd", 24. * B * d + 4. * B),
                              ("gnf_normal_ll_fwd", 4. * B * d + 8. * B), ("gnf_normal_ll_bwd", 8. * B * d + 8. * B)):
            if ktimes.get(kname):
                ms = sum(ktimes[kname]) / len(ktimes[kname])
                k4.append({"bound": "hbm", "kernel": kname, "achieved": nbytes / (ms / 1e3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                           "frac": nbytes / (ms / 1e3) / 1e9 / peak_gbs, "bytes_per_launch": nbytes, "avg_launch_ms": ms,
                           "note": "eager call bracketed by CUDA events: at this size the call is launch-latency bound (see "
                                   "profiles/r02*_k4_roofline.json for the same kernels at sizes beyond L2)" if nbytes < 64e6 else ""})
        return {"value": world * B * steps / (dev_ms / 1e3), "ms_per_step": dev_ms / steps, "k4_roofline": k4,
                "e2e": {"value": world * B * steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": B * d * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "achieved_tflops_step": flops_step / (dev_ms / steps / 1e3) / 1e12, "algorithmic_gflop_per_step": flops_step / 1e9,
                "last_loss": last, "kernel_ms": {k: sum(v) / len(v) for k, v in ktimes.items() if v},
                "nb_steps": S_, "precision": precision, "gemm_engine": gemm, "cuda_graph": bool(use_graph)}

    use_graph_any = args.cuda_graph in ("on", "auto")
    use_graph = use_graph_any
    main_res = measure(args.mode, args.precision, args.gemm, S, args.steps, args.warmup, True, use_graph)
    extra = None
    if args.mode == "train" and not args.no_eval:
        # the metric's second half: log-likelihood evaluation (UCIExperiments.py:152-162: S = nb_steps + 20), in the
        # fast mode (tensor-core UMNN forward + TF32 conditioner GEMMs, ll tolerance 2e-3)
        try:
            extra = measure("eval", "tf32", "auto-fast", S + 20, args.steps, 3, False, use_graph_any)
        except RuntimeError as err:      # e.g. integrand weights too large to stay resident in shared memory
            if "umnn tc" not in str(err):
                raise
            G.ops.enable_kernel_timing(False)
            extra = measure("eval", "strict", "auto-fast", S + 20, args.steps, 3, False, use_graph_any)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline, _ = time_cpu(cfg, B, args.mode, S, 5, 2, budget_s=25.)
    line = {"metric": metric, "value": main_res["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"], "roofline": main_res["roofline"],
            "cpu_baseline": cpu_baseline, "achieved_tflops_step": main_res["achieved_tflops_step"],
            "algorithmic_gflop_per_step": main_res["algorithmic_gflop_per_step"], "last_loss": main_res["last_loss"],
            "kernel_ms": main_res["kernel_ms"], "k4_roofline": main_res["k4_roofline"]}
    if extra is not None:
        line["eval"] = {"metric": "loglik_eval_samples_per_s", "unit": "samples/s",
                        "precision_mode": ("tf32 (tcgen05 UMNN forward + single-pass TF32 conditioner GEMMs), ll tolerance 2e-3"
                                           if extra["precision"] == "tf32" else
                                           "strict fp32 UMNN forward (integrand too wide for the resident-weight tensor-core kernel) + "
                                           "single-pass TF32 conditioner GEMMs"),
                        **{k: extra[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "achieved_tflops_step",
                                                 "algorithmic_gflop_per_step", "nb_steps", "kernel_ms", "cuda_graph")}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
