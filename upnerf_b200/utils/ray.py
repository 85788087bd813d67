"""`get_rays` / `get_ray_directions` with the reference's signatures (utils/ray.py:5-67), plus
`refine_rays`, the fused form the train step uses (pose refinement + ray casting in one kernel,
models/nerf_system.py:158-166)."""
from __future__ import annotations

import torch

from .. import _lib as L


def get_ray_directions(H, W, K):
    """Camera-space pixel directions (utils/ray.py:5-27); dataset-time, plain tensor ops."""
    dev = K.device if isinstance(K, torch.Tensor) else None
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                          torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


class _GetRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, directions, c2w):
        R = directions.shape[0]
        ctx.set_materialize_grads(False)      # an unused output arrives as None, not as a zero tensor
        rays = torch.empty(R, 8, device=directions.device, dtype=torch.float32)
        L.pose_rays_fwd(None, None, c2w, directions, None, rays)
        ctx.save_for_backward(directions, c2w)
        return rays[:, 0:3], rays[:, 3:6]

    @staticmethod
    def backward(ctx, go, gd):
        directions, c2w = ctx.saved_tensors
        if go is None and gd is None:
            return None, None
        R = directions.shape[0]
        d_rays = torch.zeros(R, 8, device=directions.device, dtype=torch.float32)
        if go is not None:
            d_rays[:, 0:3] = go
        if gd is not None:
            d_rays[:, 3:6] = gd
        d_c2w = torch.zeros_like(c2w)
        L.get_rays_bwd(c2w, directions, d_rays, d_c2w)
        return None, d_c2w


def get_rays(directions, c2w):
    """World-space ray origins and unit directions (utils/ray.py:30-67); both branches:
    one (3,4) pose for all rays, or one pose per ray."""
    if not directions.is_cuda:
        raise L.UpnerfError("get_rays: upnerf_b200 runs on CUDA tensors only (no CPU fallback)")
    d = directions.reshape(-1, 3).contiguous().float()
    batched = c2w.dim() == 3 and directions.dim() == 2 and c2w.shape[0] == directions.shape[0]
    if not batched and c2w.dim() != 2:
        raise L.UpnerfError(f"get_rays: unsupported c2w shape {tuple(c2w.shape)}")
    o, dd = _GetRaysFn.apply(d, c2w.contiguous().float())
    return o, dd


class _RefineRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, se3_table, img_idx, c2w, directions, near_far):
        R = directions.shape[0]
        rays = torch.empty(R, 8, device=directions.device, dtype=torch.float32)
        L.pose_rays_fwd(se3_table, img_idx, c2w, directions, near_far, rays)
        ctx.save_for_backward(se3_table, img_idx, c2w, directions)
        return rays

    @staticmethod
    def backward(ctx, g):
        se3_table, img_idx, c2w, directions = ctx.saved_tensors
        d_table = torch.zeros_like(se3_table)
        L.pose_rays_bwd(se3_table, img_idx, c2w, directions, g.contiguous(), d_table)
        return d_table, None, None, None, None


def refine_rays(se3_table, img_idx, c2w, directions, near_far):
    """rays[R,8] = [o, d, near, far] after composing exp(se3_table[img_idx]) with c2w; the
    gradient flows to se3_table only (dense (N_img,6), like nn.Embedding's)."""
    return _RefineRaysFn.apply(se3_table.contiguous().float(), img_idx.contiguous().long(),
                               c2w.contiguous().float(), directions.contiguous().float(),
                               near_far.contiguous().float())
