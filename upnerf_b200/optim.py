"""`FlatAdam`: the reference's Adam(eps=1e-8) (utils/optim.py:20-44, models/nerf_system.py:41-73)
over ONE flat parameter buffer, as a single `upnerf_adam_step` launch.

The reference optimiser owns ~70 tensors and, like every torch optimiser, skips a tensor whose
`.grad` is None -- which is what the heads that are out of the graph in a schedule phase have
(`rgb_share_layer` while sched_mult == 0, the candidate head once sched_mult == 1, ...): no update,
no step increment, moments untouched.  With all tensors living in one buffer that per-tensor
history is kept per SEGMENT CLASS: `segments` is a list of `(end_offset, class_key)` runs covering
the buffer, `set_live({class_key: bool})` says which classes received a gradient this step, and each
class has its own step counter for the bias corrections.  It is a `torch.optim.Optimizer`, so
`ExponentialLR` and `state_dict()` work on it unchanged.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, flat: torch.nn.Parameter, segments, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0):
        """`weight_decay` > 0 gives torch.optim.AdamW's decoupled decay (the reference's test-time
        appearance optimiser, models/nerf_system_optmize.py:61: AdamW(lr=1e-1), default decay 1e-2)."""
        if not flat.is_cuda:
            raise L.UpnerfError("FlatAdam runs on CUDA buffers only (no CPU fallback)")
        super().__init__([flat], dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        merged = []
        for end, key in segments:           # merge neighbours of the same class
            if merged and merged[-1][1] == key:
                merged[-1] = (int(end), key)
            else:
                merged.append((int(end), key))
        if not merged or merged[-1][0] != flat.numel() or len(merged) > L.ADAM_MAX_SEGMENTS:
            raise L.UpnerfError(f"FlatAdam: bad segment table ({len(merged)} segments)")
        self.segments = merged
        self.class_steps = {key: 0 for _, key in merged}
        self.live = {key: True for _, key in merged}
        st = self.state[flat]
        st["exp_avg"] = torch.zeros_like(flat.data)
        st["exp_avg_sq"] = torch.zeros_like(flat.data)

    def set_live(self, live: dict):
        for k, v in live.items():
            if k in self.live:
                self.live[k] = bool(v)

    @torch.no_grad()
    def step(self, closure=None):
        group = self.param_groups[0]
        flat = group["params"][0]
        if flat.grad is None:
            return None
        st = self.state[flat]
        b1, b2 = group["betas"]
        lr = float(group["lr"])
        for k, alive in self.live.items():
            if alive:
                self.class_steps[k] += 1
        a = L.AdamArgs()
        a.params, a.grads = flat.data_ptr(), flat.grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
        a.n, a.n_segments = flat.numel(), len(self.segments)
        for i, (end, key) in enumerate(self.segments):
            a.seg_end[i] = end
            alive = self.live[key] and self.class_steps[key] > 0
            a.seg_live[i] = int(alive)
            if alive:
                t = self.class_steps[key]
                a.seg_step_size[i] = lr / (1.0 - b1 ** t)
                a.seg_bc2_sqrt[i] = math.sqrt(1.0 - b2 ** t)
        a.beta1, a.beta2, a.eps = float(b1), float(b2), float(group["eps"])
        a.decay_mul = 1.0 - lr * float(group.get("weight_decay", 0.0))
        L.adam_step(a)
        return None

    def state_dict(self):
        """torch's optimiser state plus the per-class step counters (the bias corrections need them)."""
        sd = super().state_dict()
        sd["class_steps"] = dict(self.class_steps)
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        steps = state_dict.pop("class_steps", None)
        super().load_state_dict(state_dict)
        if steps is not None:
            for k in self.class_steps:
                self.class_steps[k] = int(steps.get(k, 0))
