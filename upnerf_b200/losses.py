"""`UPNeRFLoss` with the reference's signature and result keys (losses.py:13-64).

`UPNeRFLoss` is the plain-torch module (API parity).  `fused_tail` is what `NeRFSystem.training_step`
runs on the CUDA path: depth affine correction + every loss term + the
gradients of all of them + psnr in ONE kernel launch (`upnerf_tail_loss`, csrc/tail.cu).
"""
from __future__ import annotations

import torch
from torch import nn


class UPNeRFLoss(nn.Module):
    def __init__(self, depth_mult=1e-4, alpha_reg=1.0, encode_feat=True, fine=True):
        super().__init__()
        self.depth_mult, self.alpha_reg, self.encode_feat, self.fine = depth_mult, alpha_reg, encode_feat, fine

    def forward(self, inputs, rgb_targets, feat_targets, depth_targets, schedule_mult):
        m, out = schedule_mult, {}
        levels = (("c", "coarse"), ("f", "fine")) if self.fine else (("c", "coarse"),)
        for tag, typ in levels:
            if m < 1:
                l_depth = (inputs[f"s_depth_{typ}"] - depth_targets).abs()
                if f"t_weight_{typ}" in inputs:
                    l_depth = l_depth * (1 - inputs[f"t_weight_{typ}"].detach())
                out[f"l_depth_{tag}"] = l_depth.mean() * self.depth_mult * (1 - m)
                if self.encode_feat:
                    out[f"l_feat_{tag}"] = ((inputs[f"feat_{typ}"] - feat_targets) ** 2).mean() * (1 - m)
                else:
                    out[f"l_c_rgb_{tag}"] = ((inputs[f"c_rgb_{typ}"] - rgb_targets) ** 2).mean() * (1 - m)
            if m > 0:
                sq = (inputs[f"s_rgb_{typ}"] - rgb_targets) ** 2
                if typ == "coarse":
                    out["l_rgb_c"] = sq.mean() * m / 2
                else:
                    out["l_rgb_f"] = (sq / (2 * inputs["t_beta"] ** 2)).mean() * m
                    out["l_beta"] = torch.log(inputs["t_beta"]).mean() * m
                    out["l_alpha"] = inputs["t_alpha"].mean() * self.alpha_reg * m
        return out


_TAIL_KEYS = (("s_depth_c", "s_depth_coarse"), ("s_depth_f", "s_depth_fine"), ("t_weight_c", "t_weight_coarse"),
              ("t_weight_f", "t_weight_fine"), ("feat_c", "feat_coarse"), ("feat_f", "feat_fine"),
              ("s_rgb_c", "s_rgb_coarse"), ("s_rgb_f", "s_rgb_fine"), ("t_beta", "t_beta"), ("t_alpha", "t_alpha"))
_NO_GRAD = ("t_weight_c", "t_weight_f")     # detached in the reference (losses.py:28,47)


def fused_tail(results, batch, depth_scale, sched_mult, *, depth_mult, alpha_reg, near, far, fine, workspace,
               sched_mult_dev=None):
    """One launch for models/nerf_system.py:169-177 + losses.py:21-64 (+ backward) + psnr.

    Returns (losses[16] device tensor -- see `upnerf_tail_loss`, roots, grads): `roots[i]` is the
    tensor of `results` whose gradient (for d loss = 1) is `grads[i]`;
    d loss / d depth_scale is accumulated straight into `depth_scale.grad`.
    `sched_mult_dev` (a CUDA fp32 scalar tensor) makes the kernel read the multiplier from device memory -- a
    captured step is replayed with a new value each time; `sched_mult` then only selects the phase.
    """
    from . import _lib as L

    rgbs = batch["rgbs"]
    dev = rgbs.device
    if not rgbs.is_cuda:
        raise L.UpnerfError("fused_tail: CUDA tensors only (no CPU fallback)")
    R = rgbs.shape[0]
    a = L.TailArgs()
    a.n_rays, a.feat_dim, a.has_fine = R, batch["feats"].shape[1], int(bool(fine))
    a.sched_mult, a.depth_mult, a.alpha_reg, a.near_, a.far_ = sched_mult, depth_mult, alpha_reg, near, far
    keep = []
    if sched_mult_dev is not None:
        if not sched_mult_dev.is_cuda or sched_mult_dev.dtype != torch.float32:
            raise L.UpnerfError("fused_tail: sched_mult_dev must be a CUDA fp32 tensor")
        keep.append(sched_mult_dev)
        a.sched_mult_dev = sched_mult_dev.data_ptr()

    def f32(t):
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.contiguous().float()
        keep.append(t)
        return t.data_ptr()

    idx = batch["img_idx"]
    if not idx.is_cuda:
        raise L.UpnerfError("fused_tail: img_idx must be a CUDA tensor")
    if idx.dtype != torch.int64 or not idx.is_contiguous():
        idx = idx.contiguous().long()            # tail.cu reads int64 indices
    keep.append(idx)
    a.img_idx = idx.data_ptr()
    a.inv_depths, a.rgbs, a.feats = f32(batch["inv_depths"]), f32(rgbs), f32(batch["feats"])
    a.depth_scale = f32(depth_scale)
    if depth_scale.grad is not None:
        a.d_depth_scale = depth_scale.grad.data_ptr()
    roots, grads = [], []
    lo, hi = sched_mult < 1, sched_mult > 0
    live = {"s_depth_c": lo, "s_depth_f": lo and fine, "feat_c": lo, "feat_f": lo and fine, "s_rgb_c": hi,
            "s_rgb_f": hi and fine, "t_beta": hi and fine, "t_alpha": hi and fine}
    for field, key in _TAIL_KEYS:
        t = results.get(key)
        if t is None:
            continue
        setattr(a, field, f32(t))
        # a gradient only where a loss term of this phase reads the tensor (losses.py:21-64)
        if live.get(field, False) and t.requires_grad:
            g = torch.empty_like(t, dtype=torch.float32, memory_format=torch.contiguous_format)
            setattr(a, "g_" + field, g.data_ptr())
            roots.append(t)
            grads.append(g)
    losses = torch.empty(L.TAIL_LOSS_SLOTS, device=dev, dtype=torch.float32)
    a.losses = losses.data_ptr()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    L.tail_loss(a)
    return losses, roots, grads
