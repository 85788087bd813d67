"""`UPNeRFLoss` with the reference's signature and result keys (losses.py:13-64).

Per-ray tail of the train step, plain torch.
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
