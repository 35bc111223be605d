"""Host <-> device plumbing around the path: pinned-memory batch prefetch and pipelined scalar read-back.

The reference feeds the model from a pinned-memory DataLoader with non-blocking copies (HF Trainer,
`dataloader_pin_memory=True`; ref:run.sh / ref:src/trainer_seq2seq.py:1037 `training_step`) and reads the loss for
logging.  On a step of ~100 ms a blocking input copy plus a blocking `.item()` leave the GPU idle for several
milliseconds per step, so both are pipelined here:

  * `HostPrefetcher.stage(tensors)` enqueues the host->device copies of the NEXT batch on a dedicated copy stream while
    the current step computes; `take(handle)` makes the compute stream wait for exactly that batch.
  * `ScalarReader.submit(t)` enqueues the device->host copy of a scalar into pinned memory and returns a handle whose
    `.result()` is read one step later, when the copy has long finished.

Only streams, events and copies: no arithmetic.
"""
from typing import List, Sequence

import torch


class _Staged:
    __slots__ = ("tensors", "event")

    def __init__(self, tensors, event):
        self.tensors = tensors
        self.event = event


class HostPrefetcher:
    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPrefetcher needs a CUDA device")
        self.stream = torch.cuda.Stream(device=self.device)

    def stage(self, host_tensors: Sequence[torch.Tensor]) -> _Staged:
        """Start copying a batch (pinned host tensors) to the device on the copy stream."""
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return _Staged(dev, ev)

    def take(self, staged: _Staged) -> List[torch.Tensor]:
        """Order the current stream after the staged copies and hand the device tensors over to it."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(staged.event)
        for t in staged.tensors:
            t.record_stream(cur)      # the caching allocator must not recycle them while the compute stream reads them
        return staged.tensors


class _PendingScalar:
    __slots__ = ("buf", "event")

    def __init__(self, buf, event):
        self.buf = buf
        self.event = event

    def result(self) -> float:
        self.event.synchronize()
        return float(self.buf.item())


class ScalarReader:
    """Device scalar -> pinned host memory without stalling the launching thread until the value is actually needed."""

    def __init__(self, depth: int = 4):
        self._bufs = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._i = 0

    def submit(self, t: torch.Tensor) -> _PendingScalar:
        buf = self._bufs[self._i % len(self._bufs)]
        self._i += 1
        buf.copy_(t.detach().reshape(1).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return _PendingScalar(buf, ev)
