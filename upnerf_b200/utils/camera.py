"""`lie.se3_to_SE3` and `pose.compose` with the reference's signatures (utils/camera.py:43-58,
87-98); module-level singletons `pose` and `lie` like the reference.  Only the members on the
accelerated path are provided."""
from __future__ import annotations

import torch

from .. import _lib as L


class _Se3ExpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wu):
        out = torch.empty(wu.shape[0], 3, 4, device=wu.device, dtype=torch.float32)
        L.se3_exp_fwd(wu, out)
        ctx.save_for_backward(wu)
        return out

    @staticmethod
    def backward(ctx, g):
        (wu,) = ctx.saved_tensors
        d = torch.empty_like(wu)
        L.se3_exp_bwd(wu, g.contiguous(), d)
        return d


class _ComposeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        n = max(a.shape[0] if a.dim() == 3 else 1, b.shape[0] if b.dim() == 3 else 1)
        out = torch.empty(n, 3, 4, device=a.device, dtype=torch.float32)
        L.pose_compose_fwd(a, a.dim() == 2, b, b.dim() == 2, n, out)
        ctx.save_for_backward(a, b)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        n = ctx.n
        da = torch.empty(n, 3, 4, device=g.device) if ctx.needs_input_grad[0] else None
        db = torch.empty(n, 3, 4, device=g.device) if ctx.needs_input_grad[1] else None
        L.pose_compose_bwd(a, a.dim() == 2, b, b.dim() == 2, g.contiguous(), n, da, db)
        if da is not None and a.dim() == 2:
            da = da.sum(0)
        if db is not None and b.dim() == 2:
            db = db.sum(0)
        return da, db


def _as_poses(x):
    if not x.is_cuda:
        raise L.UpnerfError("upnerf_b200 pose ops run on CUDA tensors only (no CPU fallback)")
    return x.contiguous().float()


class Pose:
    def compose_pair(self, pose_a, pose_b):
        """pose_new(x) = pose_b o pose_a(x) (utils/camera.py:51-58); (...,3,4), broadcasting a lone (3,4)."""
        a, b = _as_poses(pose_a), _as_poses(pose_b)
        lead = a.shape[:-2] if a.dim() > 2 else b.shape[:-2]
        a2 = a.reshape(-1, 3, 4) if a.dim() > 2 else a
        b2 = b.reshape(-1, 3, 4) if b.dim() > 2 else b
        out = _ComposeFn.apply(a2, b2)
        return out.reshape(*lead, 3, 4) if len(lead) else out[0]

    def compose(self, pose_list):
        """Left fold of compose_pair (utils/camera.py:43-49)."""
        out = pose_list[0]
        for p in pose_list[1:]:
            out = self.compose_pair(out, p)
        return out


class Lie:
    def se3_to_SE3(self, wu):
        """Exponential map (...,6) -> (...,3,4) with the reference's order-10 Taylor coefficients
        (utils/camera.py:87-98,126-152)."""
        if not wu.is_cuda:
            raise L.UpnerfError("se3_to_SE3: upnerf_b200 runs on CUDA tensors only (no CPU fallback)")
        lead = wu.shape[:-1]
        out = _Se3ExpFn.apply(wu.reshape(-1, 6).contiguous().float())
        return out.reshape(*lead, 3, 4)


pose = Pose()
lie = Lie()
