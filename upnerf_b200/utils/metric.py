"""Evaluation metrics with the reference's names and signatures (utils/metric.py:10-77, helpers from
utils/camera.py:262-382) -- SURVEY.md section 8 row f4.

Per-scene, per-image quantities (763 poses, one 3x3 SVD): plain device-agnostic torch ops that run
wherever the tensors live -- on the GPU the whole pose evaluation is a handful of launches with no
host round trip until the caller reads the means.  Parity: `pose_metric`, `prealign_cameras`,
`evaluate_camera_alignment`, `psnr` are pinned by tests/golden/pose_metric.npz (written from the real
utils/metric.py).  `ssim` restates kornia.losses.ssim_loss (kornia is a third-party dependency the
reference does not pin and that is absent here: window 3, Gaussian sigma 1.5, reflect padding,
C1 = 0.01^2, C2 = 0.03^2, loss = mean(clamp((1 - ssim) / 2, 0, 1))) -- parity unpinned for `ssim`.
LPIPS needs the pretrained AlexNet weights of the `lpips` package (no network here): not provided.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mse(image_pred, image_gt, valid_mask=None, reduction="mean"):
    value = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        value = value[valid_mask]
    if reduction == "mean":
        return torch.mean(value)
    return value


def psnr(image_pred, image_gt, valid_mask=None, reduction="mean"):
    return -10 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))


def ssim(image_pred, image_gt, reduction="mean"):
    """image_pred, image_gt: (1, 3, H, W); returns 1 - 2 * dssim in [-1, 1] (utils/metric.py:24-31)."""
    win, sigma, c1, c2 = 3, 1.5, 0.01 ** 2, 0.03 ** 2
    x = torch.arange(win, dtype=image_pred.dtype, device=image_pred.device) - win // 2
    g = torch.exp(-x ** 2 / (2 * sigma ** 2))
    g = g / g.sum()
    k2 = (g[:, None] * g[None, :])[None, None].expand(image_pred.shape[1], 1, win, win)

    def filt(t):
        return F.conv2d(F.pad(t, (1, 1, 1, 1), mode="reflect"), k2, groups=t.shape[1])

    mu1, mu2 = filt(image_pred), filt(image_gt)
    s11 = filt(image_pred * image_pred) - mu1 * mu1
    s22 = filt(image_gt * image_gt) - mu2 * mu2
    s12 = filt(image_pred * image_gt) - mu1 * mu2
    ssim_map = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    loss = torch.clamp((1 - ssim_map) / 2, 0, 1)
    dssim_ = loss.mean() if reduction == "mean" else loss
    return 1 - 2 * dssim_


# ------------------------------------------------------------------ camera helpers (utils/camera.py)
def _pose(R=None, t=None):
    """utils/camera.py:12-33."""
    if R is None:
        R = torch.eye(3, device=t.device).repeat(*t.shape[:-1], 1, 1)
    elif t is None:
        t = torch.zeros(R.shape[:-1], device=R.device)
    return torch.cat([R.float(), t.float()[..., None]], dim=-1)


def _invert(pose):
    """utils/camera.py:35-41 (use_inverse=False)."""
    R, t = pose[..., :3], pose[..., 3:]
    R_inv = R.transpose(-1, -2)
    return _pose(R=R_inv, t=(-R_inv @ t)[..., 0])


def _compose_pair(pose_a, pose_b):
    """utils/camera.py:51-58."""
    R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
    R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
    return _pose(R=R_b @ R_a, t=(R_b @ t_a + t_b)[..., 0])


def cam2world(X, pose):
    """utils/camera.py:282-285."""
    X_hom = torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)
    return X_hom @ _invert(pose).transpose(-1, -2)


def rotation_distance(R1, R2, eps=1e-7):
    """utils/camera.py:354-361."""
    R_diff = R1 @ R2.transpose(-2, -1)
    trace = R_diff[..., 0, 0] + R_diff[..., 1, 1] + R_diff[..., 2, 2]
    return ((trace - 1) / 2).clamp(-1 + eps, 1 - eps).acos()


class Sim3(dict):
    __getattr__ = dict.__getitem__


def procrustes_analysis(X0, X1):
    """utils/camera.py:364-382: similarity aligning X1 to X0 (SVD in double, like the reference)."""
    t0, t1 = X0.mean(dim=0, keepdim=True), X1.mean(dim=0, keepdim=True)
    X0c, X1c = X0 - t0, X1 - t1
    s0 = (X0c ** 2).sum(dim=-1).mean().sqrt()
    s1 = (X1c ** 2).sum(dim=-1).mean().sqrt()
    U, S, Vh = torch.linalg.svd((X0c / s0).t().double() @ (X1c / s1).double(), full_matrices=False)
    R = (U @ Vh).float()
    # `if R.det() < 0: R[2] *= -1` without the host sync of a Python branch on a device scalar
    flip = torch.where(torch.linalg.det(R) < 0, -1.0, 1.0).to(R.dtype)
    R = torch.cat([R[:2], R[2:] * flip], 0)
    return Sim3(t0=t0[0], t1=t1[0], s0=s0, s1=s1, R=R)


def parse_raw_camera(pose_raw):
    """utils/metric.py:35-40 (accepts a batch: the reference maps it over the poses one by one)."""
    pose_flip = _pose(R=torch.diag(torch.tensor([1.0, -1.0, -1.0], device=pose_raw.device)))
    pose = _compose_pair(pose_flip, pose_raw[..., :3, :])
    pose = _invert(pose)
    return _compose_pair(pose_flip, pose)


def prealign_cameras(pose, pose_GT):
    """utils/metric.py:43-54."""
    pose, pose_GT = pose.float(), pose_GT.float()
    center = torch.zeros(1, 1, 3, device=pose.device)
    center_pred = cam2world(center, pose)[:, 0]
    center_GT = cam2world(center, pose_GT)[:, 0]
    sim3 = procrustes_analysis(center_GT, center_pred)
    center_aligned = (center_pred - sim3.t1) / sim3.s1 @ sim3.R.t() * sim3.s0 + sim3.t0
    R_aligned = pose[..., :3] @ sim3.R.t()
    t_aligned = (-R_aligned @ center_aligned[..., None])[..., 0]
    return _pose(R=R_aligned, t=t_aligned), sim3


def evaluate_camera_alignment(pose_aligned, pose_GT):
    """utils/metric.py:57-64."""
    R_aligned, t_aligned = pose_aligned.split([3, 1], dim=-1)
    R_GT, t_GT = pose_GT.split([3, 1], dim=-1)
    return dict(R=rotation_distance(R_aligned, R_GT), t=(t_aligned - t_GT)[..., 0].norm(dim=-1))


def pose_metric(refine_poses, gt_poses):
    """utils/metric.py:67-77: (error dict of per-image R [rad] / t, aligned poses, parsed GT poses)."""
    refine_poses = parse_raw_camera(refine_poses.float())
    gt_poses = parse_raw_camera(gt_poses.float())
    try:
        aligned_pose, _ = prealign_cameras(refine_poses, gt_poses)
        error = evaluate_camera_alignment(aligned_pose, gt_poses)
    except Exception:       # the reference's bare except: SVD did not converge
        aligned_pose, error = refine_poses, None
        print("pose alignment is not converged")
    return error, aligned_pose, gt_poses
