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
        if p.requires_grad:       # a parameter no rank used ends as zeros (group: rank-invariant tensor list; bucket: flat view)
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


def _worker_rank_dependent(rank, world, port, out):
    """A branch that only rank 0's data exercises (ranks must still issue identical collectives), then a second
    backward inside one window (must raise, not corrupt), then reduce_now() on locally accumulated gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtasr_b200.dp import GradGroupReducer
    torch.manual_seed(0)
    trunk = torch.nn.Linear(8, 8)
    branch_a, branch_b, tail = torch.nn.Linear(8, 4), torch.nn.Linear(8, 4), torch.nn.Linear(8, 8)
    params = list(tail.parameters()) + list(branch_b.parameters()) + list(branch_a.parameters()) + list(trunk.parameters())
    red = GradGroupReducer(params, group_bytes=64)            # one or two parameters per group
    assert len(red.groups) >= 4

    def loss_of(r, x):
        h = torch.relu(trunk(x))
        y = branch_a(h).pow(2).mean()
        if r == 0:                                            # data-dependent path: rank 1 never touches branch_b / tail
            y = y + branch_b(tail(h)).pow(2).mean()
        return y

    xs = [torch.randn(5, 8, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
    for p in params:
        p.grad = None
    red.begin()
    loss_of(rank, xs[rank]).backward()
    red.finish()
    ref = [torch.zeros_like(p) for p in params]
    for r in range(world):
        gs = torch.autograd.grad(loss_of(r, xs[r]), params, allow_unused=True)
        for acc, g in zip(ref, gs):
            if g is not None:
                acc += g / world
    ok = all(p.grad is not None and torch.allclose(p.grad, r_, atol=1e-6) for p, r_ in zip(params, ref))
    # second backward in the same window: loud
    raised = False
    try:
        loss_of(0, xs[rank]).backward()
    except RuntimeError as e:
        raised = "already reduced" in str(e)
    # PCGrad-style: gradients formed locally (no hooks), reduced once
    for p in params:
        p.grad = None
    gs = torch.autograd.grad(loss_of(0, xs[rank]), params)
    for p, g in zip(params, gs):
        p.grad = g * (rank + 1.0)
    red.reduce_now()
    mean_scale = sum(r + 1.0 for r in range(world)) / world
    ok2 = True
    for p in params:
        gsum = torch.zeros_like(p)
        for r in range(world):
            (g,) = torch.autograd.grad(loss_of(0, xs[r]), [p])
            gsum += g * (r + 1.0) / world
        ok2 = ok2 and torch.allclose(p.grad, gsum, atol=1e-6)
    out[rank] = (bool(ok), bool(raised), bool(ok2), mean_scale)
    dist.destroy_process_group()


def test_group_reducer_rank_dependent_graph_and_reentry_world2_gloo():
    mp.set_start_method("spawn", force=True)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker_rank_dependent, args=(2, _free_port(), out), nprocs=2, join=True)
        for r in range(2):
            assert out[r][0] is True, "reduced gradients differ from the single-process mean"
            assert out[r][1] is True, "second backward inside one window did not raise"
            assert out[r][2] is True, "reduce_now() on locally formed gradients"
