"""PCGrad over the serialized-CTC heads without the reference's extra encoder backward and host round trips.

The reference's training step (ref:src/trainer_seq2seq.py:1071-1141) back-propagates K + 1 times through encoder + separator:
once per head loss for the "shared" parameters (`model.encoder`, `model.separator`), projects every head gradient off the
ones it conflicts with -- branching on the host for each ordered head pair (`if dot < 0`, one synchronisation each) -- then
runs the full `backward(loss)` and OVERWRITES the shared parameters' gradients with the sum of the projected ones.

Same result, cheaper:
  * the full backward is restricted to the NON-shared parameters (CTC heads, decoder, projections): everything it would
    have produced for the shared parameters is discarded by the reference anyway, and autograd does not enter the
    encoder / separator when none of their parameters is requested -- K encoder backward passes instead of K + 1;
  * each head's shared gradient is written into one row of a flat (K, P) fp32 buffer; the conflict test and the projection
    run on the device (csrc/elementwise.cu pcgrad_dots / pcgrad_project: both dot products in one pass, the branch taken by
    the kernel), in the reference's order (i outer, j inner, head i's vector updated in place and the CURRENT vectors of the
    other heads used), so there is no host synchronisation in the step;
  * `.grad` of the shared parameters become views into the summed buffer.
`tests/test_pcgrad_gpu.py` checks it against a literal restatement of the reference loop.
"""
from typing import List, Optional, Sequence

import torch

from . import kernels as K


def shared_and_other_params(model):
    """The reference's split (ref:src/trainer_seq2seq.py:1082-1087): encoder + separator are shared, the rest is not."""
    shared = []
    for name in ("encoder", "separator"):
        if hasattr(model, name):
            shared += [p for p in getattr(model, name).parameters() if p.requires_grad]
    ids = {id(p) for p in shared}
    other = [p for p in model.parameters() if p.requires_grad and id(p) not in ids]
    return shared, other


def pcgrad_backward(loss: torch.Tensor, ctc_per_head: Optional[Sequence[torch.Tensor]], shared: List[torch.nn.Parameter],
                    other: List[torch.nn.Parameter], grad_accumulation_steps: int = 1) -> None:
    """Fill `.grad` of `shared` and `other` the way the reference's training step does (see module docstring)."""
    heads = [h for h in (ctc_per_head or []) if (h.mean() if h.dim() > 0 else h).requires_grad]
    if len(heads) < 2 or not shared:
        loss.backward()
        return
    scale = 1.0 / float(grad_accumulation_steps)
    sizes = [p.numel() for p in shared]
    P = sum(sizes)
    flat = torch.zeros(len(heads), P, device=shared[0].device, dtype=torch.float32)
    for i, h in enumerate(heads):
        li = (h.mean() if h.dim() > 0 else h) * scale
        gs = torch.autograd.grad(li, shared, retain_graph=True, allow_unused=True)
        off = 0
        for g, n in zip(gs, sizes):
            if g is not None:
                flat[i, off:off + n].copy_(g.reshape(-1))
            off += n
    if other:
        go = torch.autograd.grad(loss, other, allow_unused=True)      # never enters encoder / separator
        for p, g in zip(other, go):
            if g is not None:
                p.grad = g if p.grad is None else p.grad + g
    scratch = torch.empty(2, device=flat.device, dtype=torch.float32)
    Kh = len(heads)
    for i in range(Kh):
        for j in range(Kh):
            if i != j:
                K.pcgrad_project_(flat[i], flat[j], scratch)
    total = flat.sum(0)
    off = 0
    for p, n in zip(shared, sizes):
        p.grad = total[off:off + n].view_as(p)
        off += n
