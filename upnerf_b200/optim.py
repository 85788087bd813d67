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

    N_SCALARS = 2 * L.ADAM_MAX_SEGMENTS + 1      # seg_step_size | seg_bc2_sqrt | decay_mul

    def advance(self):
        """Host half of one step: bump the step counter of every live class and return the per-segment
        scalars of THIS step as (live flags, [N_SCALARS] floats: step_size | bc2_sqrt | decay_mul)."""
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        lr = float(group["lr"])
        for k, alive in self.live.items():
            if alive:
                self.class_steps[k] += 1
        M = L.ADAM_MAX_SEGMENTS
        vals = [0.0] * self.N_SCALARS
        flags = []
        for i, (_, key) in enumerate(self.segments):
            alive = self.live[key] and self.class_steps[key] > 0
            flags.append(int(alive))
            if alive:
                t = self.class_steps[key]
                vals[i] = lr / (1.0 - b1 ** t)
                vals[M + i] = math.sqrt(1.0 - b2 ** t)
        vals[2 * M] = 1.0 - lr * float(group.get("weight_decay", 0.0))
        return flags, vals

    def launch(self, flags, vals=None, dev_scalars=None):
        """Device half: one `upnerf_adam_step`.  The scalars go by value (`vals`) or -- inside a captured
        CUDA graph -- are read from `dev_scalars` ([N_SCALARS] fp32 on the device, rewritten before every replay)."""
        group = self.param_groups[0]
        flat = group["params"][0]
        st = self.state[flat]
        b1, b2 = group["betas"]
        M = L.ADAM_MAX_SEGMENTS
        a = L.AdamArgs()
        a.params, a.grads = flat.data_ptr(), flat.grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
        a.n, a.n_segments = flat.numel(), len(self.segments)
        for i, (end, _) in enumerate(self.segments):
            a.seg_end[i] = end
            a.seg_live[i] = flags[i]
            if vals is not None:
                a.seg_step_size[i] = vals[i]
                a.seg_bc2_sqrt[i] = vals[M + i]
        a.beta1, a.beta2, a.eps = float(b1), float(b2), float(group["eps"])
        if vals is not None:
            a.decay_mul = vals[2 * M]
        if dev_scalars is not None:
            if (not dev_scalars.is_cuda or dev_scalars.dtype != torch.float32 or not dev_scalars.is_contiguous()
                    or dev_scalars.numel() < self.N_SCALARS):
                raise L.UpnerfError("FlatAdam.launch: dev_scalars must be a contiguous CUDA fp32 tensor of N_SCALARS")
            a.dev_scalars = dev_scalars.data_ptr()
        L.adam_step(a)

    def live_flags(self):
        """Per-segment liveness as the NEXT `advance()` will see it (the part of a step a captured graph bakes in)."""
        return [int(self.live[key] and self.class_steps[key] + int(self.live[key]) > 0) for _, key in self.segments]

    @torch.no_grad()
    def step(self, closure=None):
        flat = self.param_groups[0]["params"][0]
        if flat.grad is None:
            return None
        flags, vals = self.advance()
        self.launch(flags, vals=vals)
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
