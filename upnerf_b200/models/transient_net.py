"""Per-ray transient MLP with the reference's state_dict (models/transient_net.py:6-38).

Adjacent to the hot path (per RAY, <0.5% of the flops).  The parameters are
plain torch modules so reference checkpoints load unchanged.  With `precision = "bf16"` on CUDA tensors
(what `NeRFSystem` sets in its bf16 mode) forward and backward run in libupnerf_b200.so
(`upnerf_tnet_fwd/bwd`: tcgen05 bf16 GEMMs, fp32 accumulation, fp32 heads); otherwise -- the fp32
validation mode and CPU construction-time use -- the torch modules evaluate it in fp32.
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


class _TnetFn(torch.autograd.Function):
    """upnerf_tnet_fwd / upnerf_tnet_bwd.  Parameter gradients are ACCUMULATED by the kernels: into each
    parameter's existing `.grad` when `sink` is set (views of NeRFSystem's flat gradient buffer; autograd then
    gets None), else into fresh zero tensors handed back to autograd."""

    @staticmethod
    def forward(ctx, net, feat, ts, sink, *params):
        from .. import _lib as L

        ctx.set_materialize_grads(False)
        R = feat.shape[0]
        dev = feat.device
        a = L.TnetArgs()
        a.n_rays, a.n_images = R, net.embedding_t.weight.shape[0]
        a.feat_dim, a.hidden, a.transient_dim = feat.shape[1], 256, net.trasient_dim
        a.beta_min = float(net.beta_min)
        feat = feat.detach().contiguous().float()
        ts = ts.contiguous().long()
        a.feats, a.img_idx = feat.data_ptr(), ts.data_ptr()
        for i, p in enumerate(params):
            a.params[i] = p.data_ptr()
        alpha = torch.empty(R, 1, device=dev)
        beta = torch.empty(R, 1, device=dev)
        rgb = torch.empty(R, 3, device=dev)
        a.alpha, a.beta, a.rgb = alpha.data_ptr(), beta.data_ptr(), rgb.data_ptr()
        ws = torch.empty(L.tnet_workspace_bytes(a), device=dev, dtype=torch.uint8)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        L.tnet_fwd(a)
        ctx.a, ctx.keep, ctx.sink, ctx.params = a, (feat, ts, ws), sink, params
        return alpha, rgb, beta

    @staticmethod
    def backward(ctx, g_alpha, g_rgb, g_beta):
        from .. import _lib as L

        a, params = ctx.a, ctx.params
        hold, grads = [], [None] * len(params)
        for name, g in (("g_alpha", g_alpha), ("g_rgb", g_rgb), ("g_beta", g_beta)):
            if g is not None:
                g = g.contiguous().float()
                hold.append(g)
            setattr(a, name, None if g is None else g.data_ptr())
        for i, p in enumerate(params):
            if not ctx.needs_input_grad[4 + i]:
                a.d_params[i] = None
            elif ctx.sink and p.grad is not None:
                a.d_params[i] = p.grad.data_ptr()
            else:
                grads[i] = torch.zeros_like(p)
                a.d_params[i] = grads[i].data_ptr()
        L.tnet_bwd(a)
        return (None, None, None, None, *grads)


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

    precision = "fp32"       # "bf16": the tensor-core path (CUDA tensors); NeRFSystem sets it from kernel.precision
    grad_sink = False        # accumulate parameter gradients straight into existing .grad buffers

    def forward(self, feat, ts):
        if self.precision == "bf16" and feat.is_cuda:
            alpha, rgb, beta = _TnetFn.apply(self, feat, ts, bool(self.grad_sink), *self.parameters())
            return {"alpha": alpha, "rgb": rgb, "beta": beta}
        enc = self.feat_encoder(feat)
        joint = self.t_encoder(torch.cat([self.final_encoder(enc), _Lookup.apply(self.embedding_t.weight, ts)], -1))
        alpha = self.alpha_layer(enc)
        return {"alpha": alpha, "rgb": self.rgb_layer(joint), "beta": self.beta_layer(joint) * alpha + self.beta_min}
