"""Gradient exchange over NVLink peer memory (csrc/peer.cu, gnf_b200.dist.PeerGroup): two ranks on two GPUs of one node.
Skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_peer.py -m gpu` runs it."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, numel, out):
    import torch.distributed as dist
    import gnf_b200 as G
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        grp = G.dist.PeerGroup(numel, dev)
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        errs = []
        for step in range(4):                                  # eager calls: the barrier epoch advances on the device
            mine = torch.randn(grp.numel, device=dev, generator=gen)
            ref = mine.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            grp.flat.copy_(mine)
            grp.allreduce_avg()
            errs.append(float((grp.flat - ref / world).abs().max()))
        # the same inside a captured CUDA graph, replayed with fresh data
        src = torch.zeros(grp.numel, device=dev)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                grp.flat.copy_(src)
                grp.allreduce_avg()
        torch.cuda.current_stream().wait_stream(s)
        for step in range(3):
            mine = torch.randn(grp.numel, device=dev, generator=gen)
            ref = mine.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            src.copy_(mine)
            g.replay()
            torch.cuda.synchronize()
            errs.append(float((grp.flat - ref / world).abs().max()))
        # bit-identical on every rank
        gathered = [torch.empty_like(grp.flat) for _ in range(world)]
        dist.all_gather(gathered, grp.flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        # GradBucket(peer=True): the training protocol lands on the same kernel
        torch.manual_seed(0)
        lin = torch.nn.Linear(37, 11).to(dev)
        G.dist.broadcast_parameters(lin)
        bucket = G.dist.GradBucket(lin.parameters(), peer=True)
        bucket.begin_step()
        x = torch.randn(8, 37, device=dev, generator=gen)
        lin(x).square().mean().backward()
        local = [p.grad.clone() for p in lin.parameters()]
        bucket.finish_step()
        gerr = 0.
        for p, l in zip(lin.parameters(), local):
            ref = l.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            gerr = max(gerr, float((p.grad - ref / world).abs().max()))
        if rank == 0:
            torch.save(dict(errs=errs, same=same, peer_active=bucket.peer is not None, gerr=gerr), out)
        dist.barrier()
        grp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("numel", [4, 1000, 950_003])
def test_peer_allreduce_two_ranks(tmp_path, numel):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), numel, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["peer_active"], "CUDA IPC peer memory could not be set up between the two GPUs"
    assert max(res["errs"]) < 1e-6, res["errs"]
    assert res["same"]
    assert res["gerr"] < 1e-6
