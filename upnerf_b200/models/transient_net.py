"""Per-ray transient MLP with the reference's state_dict (models/transient_net.py:6-38).

Adjacent to the hot path (per RAY, <0.5% of the flops): kept as
plain torch modules so reference checkpoints load unchanged.
"""
from __future__ import annotations

import torch
from torch import nn


class _Lookup(torch.autograd.Function):
    """`embedding(idx)` whose backward is one index_add_ (atomics) instead of the sort-based dense
    gradient of nn.Embedding (~20 launches): same values up to fp32 summation order."""

    @staticmethod
    def forward(ctx, weight, idx):
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return weight.index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        gw = torch.zeros(ctx.shape, device=g.device, dtype=g.dtype)
        gw.index_add_(0, idx, g)
        return gw, None


class TransientNet(nn.Module):
    def __init__(self, N_images, beta_min=0.1, trasient_dim=128, feat_dim=384):
        super().__init__()
        self.beta_min, self.trasient_dim = beta_min, trasient_dim
        self.embedding_t = nn.Embedding(N_images, trasient_dim)
        layers, n_in = [], feat_dim
        for _ in range(4):
            layers += [nn.Linear(n_in, 256), nn.ReLU()]
            n_in = 256
        self.feat_encoder = nn.Sequential(*layers)
        self.final_encoder = nn.Linear(256, 256)
        self.t_encoder = nn.Sequential(nn.Linear(256 + trasient_dim, 128), nn.ReLU())
        self.alpha_layer = nn.Sequential(nn.Linear(256, 1), nn.Sigmoid())
        self.beta_layer = nn.Sequential(nn.Linear(128, 1), nn.Softplus())
        self.rgb_layer = nn.Sequential(nn.Linear(128, 3), nn.Sigmoid())

    def forward(self, feat, ts):
        enc = self.feat_encoder(feat)
        joint = self.t_encoder(torch.cat([self.final_encoder(enc), _Lookup.apply(self.embedding_t.weight, ts)], -1))
        alpha = self.alpha_layer(enc)
        return {"alpha": alpha, "rgb": self.rgb_layer(joint), "beta": self.beta_layer(joint) * alpha + self.beta_min}
