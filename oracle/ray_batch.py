"""CPU restatement of the reference's training-ray sample path (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does).

`getitem_batch` is `default_collate([PhototourismDataset.__getitem__(i) for i in idx])` for split
"train" (datasets/phototourism.py:420-454), vectorised over the batch with the same fp32 operations
in the same order.  Pinned by tests/golden/ray_batch.npz, written by oracle/make_golden.py from the
REAL `PhototourismDataset.__getitem__` + `torch.utils.data.default_collate`.
"""
from __future__ import annotations

import torch


def getitem_batch(t: dict, idx: torch.Tensor) -> dict:
    """t: all_ray_infos (N,3), all_directions (N,3), all_rgbs (N,3), poses (N_img,3,4) and, with
    features, all_pxl_coords (N,2), feat_maps (N_img,h,h,C), all_inv_depths (N)."""
    idx = idx.long()
    img_idx = t["all_ray_infos"][idx, 2].long()                                    # :422
    out = {"ray_infos": t["all_ray_infos"][idx, :2], "directions": t["all_directions"][idx],   # :423-428
           "img_idx": img_idx, "c2w": t["poses"][img_idx].float(), "rgbs": t["all_rgbs"][idx]}
    if t.get("feat_maps") is not None:
        fm = t["feat_maps"]
        h, w = fm.shape[1], fm.shape[2]
        assert h == w                                                              # :431
        pm = t["all_pxl_coords"][idx] * (h - 1)                                    # :432
        y, x = pm[:, 0], pm[:, 1]
        f = torch.floor(pm).long()                                                 # :434
        y1, x1 = f[:, 0], f[:, 1]
        y2, x2 = torch.clamp(y1 + 1, max=h - 1), torch.clamp(x1 + 1, max=h - 1)    # :435 (h for both)
        p11, p12 = fm[img_idx, y1, x1], fm[img_idx, y1, x2]                        # :436-439
        p21, p22 = fm[img_idx, y2, x1], fm[img_idx, y2, x2]
        w11 = ((y2 - y) * (x2 - x))[:, None]                                       # :441-444
        w12 = ((y2 - y) * (x - x1))[:, None]
        w21 = ((y - y1) * (x2 - x))[:, None]
        w22 = ((y - y1) * (x - x1))[:, None]
        out["feats"] = w11 * p11 + w12 * p12 + w21 * p21 + w22 * p22               # :446-450
        out["inv_depths"] = t["all_inv_depths"][idx]                               # :451
    return out


def synth_tables(n_img: int, img_h: int, img_w: int, feat_h: int, feat_dim: int, seed: int, near=0.1, far=5.0) -> dict:
    """Synthetic per-ray tables shaped like datasets/phototourism.py:213-323 builds them: every
    image contributes img_h*img_w rays; pxl coords are linspace grids in [0,1] (so the last row and
    column sit exactly on 1.0, the border case); feature vectors are unit-norm."""
    g = torch.Generator().manual_seed(seed)
    infos, coords = [], []
    for i in range(n_img):
        n = img_h * img_w
        infos.append(torch.cat([near * torch.ones(n, 1), far * torch.ones(n, 1), i * torch.ones(n, 1)], 1))
        hp = torch.linspace(0, img_h - 1, img_h) / (img_h - 1)
        wp = torch.linspace(0, img_w - 1, img_w) / (img_w - 1)
        h_, w_ = torch.meshgrid(hp, wp, indexing="ij")
        coords.append(torch.stack((h_, w_), -1).view(-1, 2))
    N = n_img * img_h * img_w
    fm = torch.randn(n_img, feat_h, feat_h, feat_dim, generator=g)
    fm = fm / torch.norm(fm, dim=-1, keepdim=True)
    w = 0.3 * torch.randn(n_img, 3, generator=g)
    poses = torch.cat([torch.eye(3).expand(n_img, 3, 3) + 0.1 * torch.randn(n_img, 3, 3, generator=g),
                       w[:, :, None]], 2)
    return {"all_ray_infos": torch.cat(infos, 0), "all_directions": torch.randn(N, 3, generator=g),
            "all_rgbs": torch.rand(N, 3, generator=g), "all_pxl_coords": torch.cat(coords, 0),
            "all_inv_depths": 1 / far + (1 / near - 1 / far) * torch.rand(N, generator=g),
            "feat_maps": fm, "poses": poses.contiguous()}
