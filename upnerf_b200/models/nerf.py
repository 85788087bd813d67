"""`NeRF` with the reference's constructor, attributes and state_dict (models/nerf.py:5-147).

The module only OWNS parameters (fp32 masters, identical names/shapes/order, so reference
checkpoints load both ways).  The arithmetic runs in the CUDA library: `render_rays` hands
all parameters to `upnerf_render_fwd/bwd` as one flat buffer; `forward` (the per-sample
API of the reference) is served by the same dense-layer kernels through
`upnerf_b200.models._nerf_forward`.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib as L


def _act_linear(n_in: int, n_out: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(n_in, n_out), nn.ReLU(True))


class NeRF(nn.Module):
    def __init__(self, typ, D=8, W=256, skips=[4], encode_feat=True, feat_dim=384, xyz_L=10, dir_L=8,
                 appearance_dim=48, candidate_dim=16, c2f=None):
        super().__init__()
        self.typ, self.D, self.W, self.skips = typ, D, W, skips
        self.xyz_L, self.dir_L = xyz_L, dir_L
        self.in_channels_xyz, self.in_channels_dir = 6 * xyz_L + 3, 6 * dir_L + 3
        self.feat_dim, self.appearance_dim, self.candidate_dim = feat_dim, appearance_dim, candidate_dim
        self.encode_feat = encode_feat
        self.encode_appearance = appearance_dim > 0
        self.encode_candidate = candidate_dim > 0
        self.c2f = c2f
        # registration order defines the flat parameter layout the kernels read
        self.progress = nn.Parameter(torch.tensor(0.0))
        for i in range(D):
            n_in = self.in_channels_xyz if i == 0 else W + (self.in_channels_xyz if i in skips else 0)
            setattr(self, f"xyz_encoding_{i + 1}", _act_linear(n_in, W))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.share_sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        if encode_feat:
            self.feat_share_layer = nn.Linear(W, feat_dim)
        rgb_in = (feat_dim if encode_feat else W) + self.in_channels_dir + (appearance_dim if self.encode_appearance else 0)
        self.rgb_share_layer = nn.Sequential(nn.Linear(rgb_in, W // 2), nn.ReLU(True), nn.Linear(W // 2, 3), nn.Sigmoid())
        if self.encode_candidate:
            self.candidate_encoding = nn.Sequential(nn.Linear(W + candidate_dim, W // 2), nn.ReLU(True),
                                                    nn.Linear(W // 2, W // 2), nn.ReLU(True))
            self.candidate_sigma = nn.Sequential(nn.Linear(W // 2, 1), nn.Softplus())
            if encode_feat:
                self.feat_candidate_layer = nn.Linear(W // 2, feat_dim)
            else:
                self.rgb_candidate_layer = nn.Linear(W // 2, 3)

    # -- per-sample API of the reference ---------------------------------------------------
    def forward(self, inputs, sched_mult, sigma_only=False):
        from ._nerf_forward import nerf_forward

        return nerf_forward(self, inputs, sched_mult, sigma_only)

    def positional_encoding(self, input, L):
        from ._nerf_forward import positional_encoding

        return positional_encoding(self, input, L)


def net_config(model) -> L.NetConfig:
    """Kernel-side description of a (reference or upnerf_b200) NeRF module."""
    if list(model.skips) != [4] or model.D != 8 or model.W != 256:
        raise L.UpnerfError("upnerf_b200 implements the D=8, W=256, skips=[4] NeRF only; got "
                            f"D={model.D} W={model.W} skips={model.skips}")
    c2f = model.c2f
    return L.NetConfig(D=model.D, W=model.W, xyz_L=model.xyz_L, dir_L=model.dir_L,
                       encode_feat=int(bool(model.encode_feat)), feat_dim=int(model.feat_dim or 0),
                       appearance_dim=int(model.appearance_dim), candidate_dim=int(model.candidate_dim),
                       encode_appearance=int(bool(model.encode_appearance)),
                       encode_candidate=int(bool(model.encode_candidate)),
                       use_c2f=int(c2f is not None), c2f_start=float(c2f[0]) if c2f else 0.0,
                       c2f_end=float(c2f[1]) if c2f else 1.0)


def flat_parameters(model) -> torch.Tensor:
    """All parameters as ONE fp32 vector in state_dict order (autograd splits the gradient back)."""
    getter = getattr(model, "_upnerf_flat", None)
    if getter is not None:      # parameters already live in one flat buffer (NeRFSystem.model_setup)
        return getter()
    return torch.cat([p.reshape(-1) for p in model.parameters()])
