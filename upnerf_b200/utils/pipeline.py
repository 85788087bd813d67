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
    """Iterate device batches from an iterable of dicts of pinned host tensors, one batch ahead.

    The device side is TWO preallocated sets of tensors filled alternately on a copy stream (no allocation,
    no `record_stream` bookkeeping per step): the copy into a set waits -- on the copy stream, never on the
    host -- for the event that marks the end of the compute-stream work that was queued when the set was last
    handed out.  A yielded batch stays valid until the batch after the next one is requested."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher: CUDA device required")
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None, None]
        self._free = [None, None]      # compute-stream event after which set i may be overwritten
        self._n = 0
        self._next = None
        self._stage()

    def _stage(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        i = self._n & 1
        self._n += 1
        bufs = self._bufs[i]
        if bufs is None or any(k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype
                               for k, v in host.items()):
            bufs = self._bufs[i] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            self.stream.wait_stream(torch.cuda.current_stream(self.device))   # (allocated on the compute stream)
        with torch.cuda.stream(self.stream):
            if self._free[i] is not None:
                self.stream.wait_event(self._free[i])
            for k, v in host.items():
                bufs[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (i, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        i, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        # everything queued on the compute stream so far (the step that consumed the other set included) precedes
        # this event: once it has fired the other set may be refilled
        fe = torch.cuda.Event()
        fe.record(cur)
        self._free[i ^ 1] = fe
        dev = self._bufs[i]
        self._stage()                   # batch i+1 starts copying before step i is issued
        return dev


class DelayedScalar:
    """`push(t)` enqueues the D2H copy of a device scalar and returns the value pushed `depth` calls earlier (None
    until then); `drain()` returns the values still in flight, oldest first (`last()`: the newest one).  Every
    pushed scalar is read on the host exactly once; with depth > 1 the host may run that many steps ahead of the
    device, which rides out a descheduled host thread without idling the GPU."""

    def __init__(self, depth: int = 1):
        self.depth = max(1, int(depth))
        self.slots = [torch.empty((), pin_memory=True) for _ in range(self.depth + 1)]
        self.events = [None] * (self.depth + 1)
        self.n = 0
        self.read = 0

    def push(self, t: torch.Tensor):
        i = self.n % (self.depth + 1)
        self.slots[i].copy_(t.detach().reshape(()), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev
        self.n += 1
        return self._read_next() if self.n - self.read > self.depth else None

    def _read_next(self):
        i = self.read % (self.depth + 1)
        self.events[i].synchronize()
        self.read += 1
        return float(self.slots[i])

    def drain(self):
        out = []
        while self.read < self.n:
            out.append(self._read_next())
        return out

    def last(self):
        vals = self.drain()
        return vals[-1] if vals else None
