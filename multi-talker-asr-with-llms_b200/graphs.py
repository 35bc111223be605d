"""One training step (forward + backward, optionally the data-parallel gradient reduction) as ONE CUDA graph.

A cfg2 step enqueues ~1750 kernels (1350 of this package through ctypes + autograd / torch glue) from a single Python
thread: ~55 us of host work per launch is as long as the GPU needs to execute them, so the eager step is bound by the HOST
(`host_issue_ms_per_step` in bench.py equals `ms_per_step`) and kernel-side gains do not show.  The step has no host
synchronisation (lengths stay on the device, the loss is read after the step), fixed shapes per bucket of the data loader
(`--group_by_length`, ref:run.sh:245) and no data-dependent control flow, so it can be captured once and replayed:

    step = GraphedTrainStep(lambda w, m, y0, y1, n0, n1: model(w, attention_mask=m, label_spks=[y0, y1],
                                                               label_spks_lengths=[n0, n1]), example_tensors, params)
    loss = step(w, m, y0, y1, n0, n1)       # copies into the static inputs, replays, returns the static loss tensor
    optimizer.step()                        # gradients are in p.grad (static tensors, rewritten by every replay and
                                            # re-attached if the caller cleared them with zero_grad(set_to_none=True))

What capture changes, and how it is kept correct:
  * parameter-derived operands (bf16 copies, the fused QKV weight) are normally cached per parameter version OUTSIDE the
    step; a replay cannot see a version bump, so during capture the caches are bypassed (`ops.capturing()`): the casts become
    graph nodes and every replay reads the current weights (0.7 ms per cfg2 step, the price an eager step also pays once
    the optimizer has touched the weights);
  * dropout seeds come from torch's CUDA generator through a torch op, which `torch.cuda.graph` registers and advances per
    replay: every replay draws fresh masks;
  * `.grad` is set to None before capture, so the backward's first write of each gradient is an assignment of a tensor from
    the graph's private pool: replays overwrite, never accumulate;
  * the gradient reducer (dp.GradGroupReducer) launches its NCCL calls on a side stream joined to the capturing stream by
    events; they are captured as part of the same graph (capture_error_mode="thread_local": the NCCL watchdog thread's CUDA
    calls do not interfere).
A new input shape needs a new capture (one `GraphedTrainStep` per shape bucket).

Requirement inherited from autograd: no autograd graph that reaches these parameters may still be alive when the step is
built.  A parameter's gradient accumulator is created with the first graph that uses it, is bound to the stream that graph
was built on, and is shared by later graphs for as long as any of them lives; an accumulator left over from an eager step on
the default stream would make the legacy stream wait on the capturing stream (cudaErrorStreamCaptureImplicit).  The usual
holder is `HybridLoss.last_ctc_per_head` (per-head losses WITH graph, kept for PCGrad): pass `release=` (e.g.
`SerializedCTCPath.release_graph`) and it is called before the warm-up.
"""
from typing import Callable, Iterable, Optional, Sequence

import torch

from . import kernels as K
from . import ops


class GraphedTrainStep:
    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], params: Iterable[torch.nn.Parameter],
                 reducer=None, backward_sm_budget: int = 0, warmup: int = 3, release: Optional[Callable[[], None]] = None):
        if release is not None:
            release()
        self.fn = fn
        self.params = [p for p in params if p.requires_grad]
        self.reducer = reducer
        self.budget = int(backward_sm_budget)
        self.static_in = [t.clone() for t in example_inputs]
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture: autotuning-free, but allocator / caches
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        l0 = K.launch_count()
        with ops.capturing():
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.static_loss = self._body()
        self.launches_per_replay = K.launch_count() - l0    # kernels of this package inside one replay
        # the tensors every replay writes the gradients into: re-attached after each replay, so a caller that clears
        # gradients between steps (`optimizer.zero_grad()` / `model.zero_grad()` set `.grad = None` by default, as the HF
        # trainer does, ref:src/trainer_seq2seq.py:1037-1141) still finds them in `.grad` after the next step
        self.static_grads = [p.grad for p in self.params]

    def _body(self) -> torch.Tensor:
        for p in self.params:
            p.grad = None
        if self.reducer is not None:
            self.reducer.begin()
        loss = self.fn(*self.static_in)
        if self.budget:
            K.set_sm_budget(self.budget)                    # the backward's persistent kernels leave SMs to NCCL
        try:
            loss.backward()
        finally:
            if self.budget:
                K.set_sm_budget(0)
        if self.reducer is not None:
            self.reducer.finish()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if len(inputs) != len(self.static_in):
            raise ValueError(f"GraphedTrainStep: expected {len(self.static_in)} tensors, got {len(inputs)}")
        for dst, src in zip(self.static_in, inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError(f"GraphedTrainStep: input {tuple(src.shape)} {src.dtype} does not match the captured "
                                 f"{tuple(dst.shape)} {dst.dtype}; capture one step per shape bucket")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.static_grads):
            if p.grad is not g:
                p.grad = g
        return self.static_loss
