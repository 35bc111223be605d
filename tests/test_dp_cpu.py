"""world_size-2 gloo test of the data-parallel gradient reduction (mtasr_b200.dp.GradBucketReducer)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, kind="bucket"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtasr_b200.dp import GradBucketReducer, GradGroupReducer
    torch.manual_seed(0)                                      # same weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    net[3].weight.requires_grad_(False)                       # frozen parameter
    unused = torch.nn.Parameter(torch.ones(5))                # trainable but never reached by the loss
    params = list(net.parameters()) + [unused]
    if kind == "bucket":
        red = GradBucketReducer(params, bucket_bytes=1024)    # tiny buckets -> several collectives
        assert len(red.buckets) > 1
    else:
        red = GradGroupReducer(params, group_bytes=1024)      # in-place, copy-free groups
        assert len(red.groups) > 1
    for step in range(2):                                     # reuse across steps
        g = torch.Generator().manual_seed(100 + rank + 10 * step)
        x = torch.randn(6, 16, generator=g)                   # each rank draws its own utterances
        if kind == "bucket":
            red.zero_grad()
        else:
            for p in params:
                p.grad = None
            red.begin()
        net(x).pow(2).mean().backward()
        red.finish()
    # single-process reference: mean of the two ranks' gradients of the last step
    ref = [torch.zeros_like(p) for p in params]
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r + 10)
        x = torch.randn(6, 16, generator=g)
        gs = torch.autograd.grad(net(x).pow(2).mean(), [p for p in params[:-1] if p.requires_grad])
        it = iter(gs)
        for i, p in enumerate(params[:-1]):
            if p.requires_grad:
                ref[i] += next(it) / world
    ok = True
    for p, r_ in zip(params, ref):
        if p.requires_grad and (kind == "bucket" or p is not unused):
            ok = ok and p.grad is not None and torch.allclose(p.grad, r_, atol=1e-6)
        else:
            ok = ok and p.grad is None
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    mp.set_start_method("spawn", force=True)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        assert out[0] is True and out[1] is True


def test_inplace_group_allreduce_world2_gloo():
    mp.set_start_method("spawn", force=True)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, _free_port(), out, "group"), nprocs=2, join=True)
        assert out[0] is True and out[1] is True
