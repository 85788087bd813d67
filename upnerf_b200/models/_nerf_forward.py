"""Per-sample `NeRF.forward` / `NeRF.positional_encoding` (reference models/nerf.py:80-147).

`render_rays` never calls these (it hands whole ray batches to `upnerf_render_fwd`, where the
heads are restructured per ray); they exist so code that uses the reference's per-sample API
keeps working.  Every layer is one launch of the library's dense kernel (`upnerf_gemm_f32`,
bias + activation in its epilogue) and the encoding is `upnerf_posenc_fwd`.  Forward only: the
outputs carry no autograd graph (training goes through `render_rays`, whose backward is
`upnerf_render_bwd`).  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib as L


def _band_weights(model, n_bands: int, device) -> torch.Tensor:
    out = torch.empty(16, device=device, dtype=torch.float32)
    c2f = model.c2f
    L.c2f_weights(model.progress.data.reshape(1), float(c2f[0]) if c2f else 0.0, float(c2f[1]) if c2f else 1.0,
                  c2f is not None, n_bands, out)
    return out


def positional_encoding(model, input: torch.Tensor, n_bands: int) -> torch.Tensor:
    """[..., N] -> [..., N + 2*N*L], layout [x | per coordinate: sin block, cos block] (nerf.py:126-147)."""
    if not input.is_cuda:
        raise L.UpnerfError("NeRF.positional_encoding: CUDA tensors only (no CPU fallback)")
    shape = input.shape
    n = shape[-1]
    if n != 3:
        raise L.UpnerfError(f"positional_encoding kernel encodes 3-vectors, got last dim {n}")
    x = input.detach().reshape(-1, n).contiguous().float()
    M = x.shape[0]
    width = n + 2 * n * n_bands
    out = torch.empty(M, width, device=x.device, dtype=torch.float32)
    L.posenc_fwd(x, x.stride(0), M, n_bands, _band_weights(model, n_bands, x.device), out, width, width, L.F32)
    return out.view(*shape[:-1], width)


def _linear(x: torch.Tensor, lin: torch.nn.Linear, act: int = 0) -> torch.Tensor:
    """y = act(x W^T + b): one dense-kernel launch."""
    M, K = x.shape
    N = lin.out_features
    w = lin.weight.data
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    ep = L.make_epilogue(bias=lin.bias.data, act=act)
    L.gemm_f32(x, (x.stride(0), x.stride(1)), w, (w.stride(0), w.stride(1)), y, (N, 1), M, N, K, ep=ep)
    return y


def _softplus(x):   # nn.Softplus(beta=1, threshold=20) of the sigma heads (nerf.py:51,74)
    return torch.nn.functional.softplus(x)


@torch.no_grad()
def nerf_forward(model, inputs: dict, sched_mult, sigma_only: bool = False) -> dict:
    ret = {}
    xyz = inputs["input_xyz"]
    if not xyz.is_cuda:
        raise L.UpnerfError("NeRF.forward: CUDA tensors only (no CPU fallback)")
    pe = positional_encoding(model, xyz, model.xyz_L)
    h = pe
    for i in range(model.D):
        if i in model.skips:
            h = torch.cat([pe, h], 1)
        h = _linear(h, getattr(model, f"xyz_encoding_{i + 1}")[0], act=1)
    ret["s_sigma"] = _softplus(_linear(h, model.share_sigma[0]))
    if sigma_only:
        return ret
    hf = _linear(h, model.xyz_encoding_final)

    def rgb_head(front):
        parts = [front, positional_encoding(model, inputs["input_dir"], model.dir_L)]
        if model.encode_appearance:
            parts.append(inputs["input_a"].float())
        q = _linear(torch.cat(parts, 1), model.rgb_share_layer[0], act=1)
        return torch.sigmoid(_linear(q, model.rgb_share_layer[2]))

    def candidate_trunk():
        g = torch.cat([hf, inputs["input_c"].float()], 1)
        g = _linear(g, model.candidate_encoding[0], act=1)
        return _linear(g, model.candidate_encoding[2], act=1)

    if model.encode_feat:
        ret["s_feat"] = _linear(hf, model.feat_share_layer)
        if sched_mult < 1 and model.encode_candidate:
            g = candidate_trunk()
            ret["c_sigma"] = _softplus(_linear(g, model.candidate_sigma[0]))
            ret["c_feat"] = _linear(g, model.feat_candidate_layer)
        if sched_mult > 0:
            ret["s_rgb"] = rgb_head(ret["s_feat"])
    else:
        ret["s_rgb"] = rgb_head(hf)
        if sched_mult < 1:
            g = candidate_trunk()
            ret["c_sigma"] = _softplus(_linear(g, model.candidate_sigma[0]))
            ret["c_rgb"] = _linear(g, model.rgb_candidate_layer)
    return ret
