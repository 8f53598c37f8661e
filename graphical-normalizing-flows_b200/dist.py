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


class _DevArray:
    """__cuda_array_interface__ view of a raw device allocation (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class PeerGroup:
    """NVLink peer-memory exchange group of the ranks of one node (csrc/peer.cu): every rank owns a cudaMalloc'ed fp32 bucket and
    a flag block, opened by all other ranks through CUDA IPC handles (exchanged once with all_gather_object).  `flat` is this
    rank's bucket as a torch tensor; `allreduce_avg()` enqueues ONE kernel that averages it over the ranks in place.
    Raises RuntimeError when peer memory cannot be set up (the caller falls back to NCCL)."""

    def __init__(self, numel, device):
        import ctypes as C
        from . import _lib as L
        lib = L.lib()
        self._lib, self._C = lib, C
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.numel = (int(numel) + 3) // 4 * 4
        self.device = torch.device(device)
        self._own, self._opened = [], []
        with torch.cuda.device(self.device):
            buf, flag = C.c_void_p(), C.c_void_p()
            ok = lib.gnf_peer_alloc(self.numel * 4, C.byref(buf)) == 0 and lib.gnf_peer_alloc(lib.gnf_peer_flag_bytes(), C.byref(flag)) == 0
            self._own = [buf, flag]
            hb, hf = C.create_string_buffer(64), C.create_string_buffer(64)
            ok = ok and lib.gnf_peer_export(buf, hb) == 0 and lib.gnf_peer_export(flag, hf) == 0
            mine = (bytes(hb.raw), bytes(hf.raw)) if ok else None
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            if any(h is None for h in everyone):
                self.close()
                raise RuntimeError("peer memory: an export failed on some rank: " + lib.gnf_last_error().decode())
            bufs, flags = (C.c_void_p * self.world)(), (C.c_void_p * self.world)()
            fail = None
            for r, (b, f) in enumerate(everyone):
                if r == self.rank:
                    bufs[r], flags[r] = buf.value, flag.value
                    continue
                pb, pf = C.c_void_p(), C.c_void_p()
                if lib.gnf_peer_import(b, C.byref(pb)) != 0 or lib.gnf_peer_import(f, C.byref(pf)) != 0:
                    fail = lib.gnf_last_error().decode()
                    break
                self._opened += [pb, pf]
                bufs[r], flags[r] = pb.value, pf.value
            status = [None] * self.world
            dist.all_gather_object(status, fail)
            if any(st is not None for st in status):
                self.close()
                raise RuntimeError("peer memory: " + "; ".join(st for st in status if st))
            self._bufs, self._flags = bufs, flags
            self.flat = torch.as_tensor(_DevArray(buf.value, self.numel), device=self.device)
        dist.barrier()

    def allreduce_avg(self):
        from . import _lib as L
        from . import ops
        ops._call("gnf_peer_allreduce_avg", self._bufs, self._flags, self.rank, self.world, self.numel, L.stream_ptr())
        ops._count()

    def close(self):
        for p in self._opened:
            self._lib.gnf_peer_close(p)
        for p in self._own:
            if p and p.value:
                self._lib.gnf_peer_free(p)
        self._opened, self._own = [], []


class GradBucket:
    """Flat gradient buffer; every ``p.grad`` aliases a slice of it after the all-reduce.

    Overlap (the single all-reduce issued after the last wgrad cost 7 % at 2..8 GPUs): the parameters are grouped
    into sub-buckets in REVERSE registration order -- the order in which the backward produces their gradients (normalizer,
    then the conditioner's layers from the last to the first, then A) -- and a post-accumulate-grad hook on the last-produced
    parameter of each sub-bucket hands it to a communication stream: that stream waits for the gradient kernels, packs the
    sub-bucket (one multi-tensor copy) and issues its NCCL all-reduce while the main stream keeps running the rest of the
    backward.  Only the last, small sub-bucket (first layer + A) is exposed.  Works inside CUDA-graph capture (the side stream
    is forked from and joined to the capturing stream)."""

    def __init__(self, params, bucket_bytes=1 << 21, overlap=True, peer=False):
        """peer=True (CUDA, world_size > 1, all ranks on one node): the bucket lives in NVLink peer memory and the average is ONE
        kernel of libgnf (PeerGroup / gnf_peer_allreduce_avg) instead of a NCCL all-reduce; silently falls back to NCCL when
        CUDA IPC is not available.  `self.peer` tells which one is active."""
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.peer = None
        if peer and is_dist() and dev.type == "cuda" and dt == torch.float32:
            try:
                self.peer = PeerGroup(self.numel, dev)
                overlap = False
            except RuntimeError as err:
                import warnings
                warnings.warn(f"GradBucket: {err}; using the NCCL all-reduce")
        self.flat = self.peer.flat[:self.numel] if self.peer is not None else torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        self._offsets = {}
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._offsets[id(p)] = (off, n)
            off += n
        # sub-buckets: contiguous runs of the parameter list, walked from the end (gradient readiness order)
        self.groups, cur, cur_bytes = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self.groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.groups.append(cur)
        self.overlap = bool(overlap) and self.flat.is_cuda
        self._comm = torch.cuda.Stream(device=dev) if self.overlap else None
        self._works, self._pending, self._armed = [], None, False
        self._group_of = {id(p): gi for gi, g in enumerate(self.groups) for p in g}
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def zero(self):
        self.flat.zero_()

    # ---- step protocol without accumulation kernels -------------------------------------------------------------
    # With p.grad aliasing the bucket, autograd ADDS every incoming gradient into it: one zero fill + one elementwise add per
    # parameter tensor and step (16 launches, 2.5 % of the cfg4 step).  begin_step() drops the gradients instead, so that
    # autograd simply adopts the tensors the backward kernels produced; the sub-buckets are packed into the flat buffer only
    # when there is something to all-reduce (one multi-tensor copy each), and finish_step() leaves p.grad pointing at the
    # averaged slices.
    def begin_step(self):
        for p in self.params:
            p.grad = None
        self._works = []
        self._pending = [len(g) for g in self.groups]
        self._armed = self.overlap and is_dist()

    def _group_slice(self, gi):
        g = self.groups[gi]
        lo = min(self._offsets[id(p)][0] for p in g)
        hi = max(self._offsets[id(p)][0] + self._offsets[id(p)][1] for p in g)
        return self.flat[lo:hi]

    def _view(self, p):
        off, n = self._offsets[id(p)]
        return self.flat[off:off + n].view_as(p)

    def _launch_group(self, gi):
        """Pack sub-bucket gi and all-reduce it on the communication stream (called when its last gradient exists)."""
        g = self.groups[gi]
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)                      # the gradient kernels enqueued so far
        with torch.cuda.stream(self._comm):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in g]
            torch._foreach_copy_([self._view(p) for p in g], grads)
            self._works.append(dist.all_reduce(self._group_slice(gi), op=dist.ReduceOp.AVG, async_op=True))

    def _on_grad(self, p):
        if not self._armed:
            return
        gi = self._group_of[id(p)]
        self._pending[gi] -= 1
        if self._pending[gi] == 0:
            self._launch_group(gi)

    def finish_step(self):
        if not is_dist():
            return
        if self._armed:
            for gi, left in enumerate(self._pending):     # parameters that received no gradient this step
                if left > 0:
                    self._pending[gi] = 0
                    self._launch_group(gi)
            for w in self._works:
                w.wait()
            torch.cuda.current_stream().wait_stream(self._comm)
            self._armed = False
        else:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            torch._foreach_copy_(self._views(), grads)
            self.allreduce_mean()
        for p in self.params:
            p.grad = self._view(p)

    def _views(self):
        return [self._view(p) for p in self.params]

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
        if self.peer is not None:
            self.peer.allreduce_avg()
        elif self.flat.is_cuda:
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
