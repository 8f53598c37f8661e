"""Data-parallel training plumbing: one process per GPU (torchrun), batch sharded on dim 0, parameters
replicated, ONE flat fp32 gradient bucket all-reduced per step.

Replaces the reference's single-process ``nn.DataParallel`` (ImageExperiments.py:168), which re-broadcasts
every parameter on every forward.  Here parameters are broadcast once; ``p.grad`` of every parameter is a
view into one contiguous buffer, so the backward kernels write (accumulate) straight into the bucket and the
collective is a single NCCL all-reduce with no packing copies.
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(model, src=0):
    """Make every rank's parameters and buffers identical to rank `src`'s (once, at start)."""
    if not is_dist():
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


class GradBucket:
    """Flat gradient buffer; every ``p.grad`` aliases a slice of it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    # ---- step protocol without accumulation kernels -------------------------------------------------------------
    # With p.grad aliasing the bucket, autograd ADDS every incoming gradient into it: one zero fill + one elementwise add per
    # parameter tensor and step (16 launches, 2.5 % of the cfg4 step).  begin_step() drops the gradients instead, so that
    # autograd simply adopts the tensors the backward kernels produced; finish_step() packs them into the flat bucket only
    # when there is something to all-reduce (one multi-tensor copy), and leaves p.grad pointing at the averaged slices.
    def begin_step(self):
        for p in self.params:
            p.grad = None

    def finish_step(self):
        if not is_dist():
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        views = self._views()
        torch._foreach_copy_(views, grads)
        self.allreduce_mean()
        for p, v in zip(self.params, views):
            p.grad = v

    def _views(self):
        out, off = [], 0
        for p in self.params:
            n = p.numel()
            out.append(self.flat[off:off + n].view_as(p))
            off += n
        return out

    def check_aliasing(self):
        """True while every p.grad still lives inside the bucket (optimizers with set_to_none break it)."""
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * self.flat.element_size()
        return all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params)

    def allreduce_mean(self):
        """Average the bucket over ranks: the gradient of  constraintsLoss - mean_global(ll)  when every rank
        holds an equal share of the batch (the constraint term is identical on every rank, so its average is
        itself)."""
        if not is_dist():
            return
        if self.flat.is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                                                     # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())


def shard_batch(x, rank=None, world=None):
    """Contiguous split of the batch on dim 0 (what nn.DataParallel.scatter does)."""
    if rank is None:
        rank = dist.get_rank() if is_dist() else 0
    if world is None:
        world = dist.get_world_size() if is_dist() else 1
    if x.shape[0] % world != 0:
        # unequal (or empty) shards would bias the AVG all-reduce of per-rank mean gradients and turn an empty shard's
        # mean log-likelihood into NaN on every rank
        raise ValueError(f"batch of {x.shape[0]} samples does not split evenly over {world} ranks: drop or pad the remainder")
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]


def decorrelate_gate_noise(model, rank=None):
    """Give every rank its own Philox counter range for the DAG gate noise."""
    if rank is None:
        rank = dist.get_rank() if is_dist() else 0
    for c in model.getConditioners():
        if hasattr(c, "_noise_rank"):
            c._noise_rank = int(rank)
