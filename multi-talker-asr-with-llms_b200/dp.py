"""Data-parallel gradient reduction for the path: the path shards by utterance, one process per GPU, and
the only collective is the gradient all-reduce (mean) of the trainable encoder / separator / CTC-head parameters --
what the reference gets from DDP inside `accelerator.backward` (ref:src/trainer_seq2seq.py:1134).

`GradBucketReducer` keeps every trainable parameter's `.grad` as a view into a few large flat fp32 buckets laid out in
reverse registration order (~ the order backward produces them: CTC heads, separator, encoder layers 23..0).  A
post-accumulate-grad hook counts arrivals per bucket and launches the bucket's asynchronous all-reduce (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) as soon as its last gradient has landed, so the wire time hides
under the rest of the backward pass.  `finish()` flushes buckets that never filled (unused parameters), waits and
turns the sums into means.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 256 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        order = list(reversed(self.params))
        self.buckets: List[dict] = []
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._works: List = []

    def _close(self, plist):
        dev = plist[0].device
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        b = dict(flat=flat, params=plist, pending=len(plist), launched=False)
        off = 0
        for p in plist:
            p._mtasr_bucket = len(self.buckets)
            p._mtasr_view = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append(b)

    def zero_grad(self):
        """Point every .grad at its bucket view and zero the buckets (call before each backward)."""
        self._works = []
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["launched"] = False
            for p in b["params"]:
                p.grad = p._mtasr_view

    def _launch(self, b):
        b["launched"] = True
        if self.world > 1:
            self._works.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        b = self.buckets[p._mtasr_bucket]
        if p.grad is not p._mtasr_view:               # autograd replaced the tensor (first accumulation): copy in
            p._mtasr_view.copy_(p.grad)
            p.grad = p._mtasr_view
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def finish(self):
        """Flush never-filled buckets, wait for the collectives and average."""
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        for w in self._works:
            w.wait()
        if self.world > 1:
            for b in self.buckets:
                b["flat"].div_(self.world)
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()


class GradGroupReducer:
    """In-place, copy-free variant: gradients stay where the backward kernels wrote them.

    DistributedDataParallel (and `GradBucketReducer`) move every gradient into a flat bucket and pre-divide it: two extra
    passes over the 2.4 GB of fp32 gradients of cfg2 plus ~500 small copy launches from autograd hooks, measured at
    3.4 ms of a 100 ms step before a single byte crosses NVLink.  Here the parameters are only *grouped* (reverse
    registration order, ~`group_bytes` each); when the last gradient of a group has landed, one coalesced NCCL call
    (`ncclGroupStart/End` around per-tensor all-reduces, `ReduceOp.AVG`) reduces the tensors in place on a side stream
    ordered after the producing kernels by an event.  `finish()` flushes groups that never completed (unused / frozen
    parameters) and makes the current stream wait for the collectives.  Call `begin()` before each backward (with
    `.grad = None`, as `optimizer.zero_grad(set_to_none=True)` leaves them).

    The sequence of collectives is RANK-INVARIANT by construction, whatever each rank's data did to its autograd graph:
      * every group is reduced with the full, fixed tensor list of its parameters -- a parameter that received no gradient
        on this rank (an unused head or branch, a data-dependent path) contributes zeros, it is never dropped from the call;
      * groups are launched strictly in index order: a group that is complete on this rank but follows an incomplete one
        waits for `finish()`, so two ranks can never issue the same collectives in different orders.
    (A parameter unused on every rank therefore ends with a zero gradient, not None, when world_size > 1.)

    One `begin()` ... `finish()` window covers exactly ONE backward pass through `.backward()`.  A gradient hook firing for
    a group that was already reduced -- a second backward (PCGrad's K+1 passes, gradient accumulation) without `begin()` --
    raises instead of silently accumulating into averaged tensors.  For those schemes accumulate locally
    (torch.autograd.grad, or several backward passes outside the window) and call `reduce_now()` once."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group_bytes: int = 256 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.groups: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > group_bytes:
                self.groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.groups.append(cur)
        self._gid = {}
        for gi, g in enumerate(self.groups):
            for p in g:
                self._gid[p] = gi
        self._cuda = bool(self.params) and self.params[0].is_cuda
        self._stream = torch.cuda.Stream(device=self.params[0].device) if self._cuda else None
        backend = dist.get_backend(group) if dist.is_initialized() else ""
        self._avg = backend == "nccl"                 # gloo has no AVG: SUM, then divide
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._works: List = []
        self.begin()

    def begin(self):
        self._pending = [len(g) for g in self.groups]
        self._launched = [False] * len(self.groups)
        self._next = 0                                # groups [0, _next) have been launched
        self._works = []

    def _reduce(self, tensors):
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        cm = getattr(dist, "_coalescing_manager", None)
        if self._avg and cm is not None and len(tensors) > 1:
            with cm(group=self.group, device=tensors[0].device, async_ops=True) as work:
                for t in tensors:
                    dist.all_reduce(t, op=op, group=self.group)
            self._works.append(work)
        else:
            for t in tensors:
                self._works.append(dist.all_reduce(t, op=op, group=self.group, async_op=True))

    def _launch(self, gi):
        self._launched[gi] = True
        if self.world == 1:
            return
        for p in self.groups[gi]:                     # rank-invariant tensor list: zeros stand in for a missing gradient
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        tensors = [p.grad for p in self.groups[gi]]
        if self._cuda:
            ev = torch.cuda.Event()
            ev.record()                               # the gradients of this group are complete on the current stream
            self._stream.wait_event(ev)
            with torch.cuda.stream(self._stream):
                self._reduce(tensors)
        else:
            self._reduce(tensors)

    def _advance(self, force: bool = False):
        """Launch, in index order, every group that is complete (or every remaining group when `force`)."""
        while self._next < len(self.groups) and (force or self._pending[self._next] == 0):
            self._launch(self._next)
            self._next += 1

    def _on_grad(self, p):
        gi = self._gid[p]
        if self._launched[gi] or self._pending[gi] <= 0:
            raise RuntimeError(
                "GradGroupReducer: a gradient arrived for a parameter whose group was already reduced in this begin()/finish() "
                "window (second backward pass without begin()?).  Reduce once per window: call begin() before every "
                ".backward(), or accumulate locally and call reduce_now().")
        self._pending[gi] -= 1
        if gi == self._next and self._pending[gi] == 0:
            self._advance()

    def finish(self):
        self._advance(force=True)
        for w in self._works:
            if w is not None:
                w.wait()
        if self._cuda:
            torch.cuda.current_stream().wait_stream(self._stream)
        if self.world > 1 and not self._avg:
            for g in self.groups:
                for p in g:
                    p.grad.div_(self.world)
        self._works = []

    def reduce_now(self):
        """Average whatever is in `.grad` right now (gradients produced outside a hook-driven backward: PCGrad's projected
        gradients, locally accumulated micro-batches)."""
        self.begin()
        self.finish()

    def remove(self):
        for h in self._hooks:
            h.remove()
