"""Host<->device plumbing around `NeRFSystem.training_step` for callers that feed HOST batches.

The reference hands each collated batch to Lightning, which copies it to the GPU and logs the loss
asynchronously (models/nerf_system.py:150-229, train.py:60-80).  With the step at ~7 ms a blocking
copy + a blocking loss read per step cost ~10 % of the step; these two helpers keep both inside the
step's shadow without changing what is computed:

  * `DevicePrefetcher`  stages batch i+1 (pinned host tensors) on a copy stream while step i runs;
  * `DelayedScalar`     reads the scalar of step i-1 while step i is in flight (one D2H per step).
"""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Iterate device batches from an iterable of dicts of pinned host tensors, one batch ahead."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher: CUDA device required")
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._stage()

    def _stage(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for v in dev.values():
            v.record_stream(cur)        # allocated on the copy stream, consumed on the compute stream
        self._stage()                   # batch i+1 starts copying before step i is issued
        return dev


class DelayedScalar:
    """`push(t)` enqueues the D2H copy of a device scalar and returns the value pushed one call
    earlier (None the first time); `last()` drains the final one."""

    def __init__(self):
        self.slots = [torch.empty((), pin_memory=True), torch.empty((), pin_memory=True)]
        self.events = [None, None]
        self.n = 0

    def push(self, t: torch.Tensor):
        i = self.n & 1
        self.slots[i].copy_(t.detach().reshape(()), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev
        self.n += 1
        return self._read(i ^ 1) if self.n > 1 else None

    def _read(self, i):
        self.events[i].synchronize()
        return float(self.slots[i])

    def last(self):
        return self._read((self.n - 1) & 1) if self.n else None
