"""GPU: the resident ray batcher against the golden fixture written from
the real `PhototourismDataset.__getitem__` and against the oracle at the reference's sizes.  Gathers and
the 4-tap interpolation are bit-exact by construction (same fp32 operations, no FMA contraction)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ray_batch as RB

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _batcher(tables, dev, **kw):
    from upnerf_b200.datasets import RayBatcher

    return RayBatcher(tables["all_ray_infos"], tables["all_directions"], tables["all_rgbs"], tables["poses"],
                      all_pxl_coords=tables.get("all_pxl_coords"), feat_maps=tables.get("feat_maps"),
                      all_inv_depths=tables.get("all_inv_depths"), device=dev, **kw)


def _same(got, want):
    assert set(got) == set(want)
    for k, v in want.items():
        g = got[k].cpu()
        assert g.dtype == v.dtype and g.shape == v.shape, k
        assert torch.equal(g, v), (k, float((g.double() - v.double()).abs().max()))


def test_golden(cuda_dev):
    z = np.load(GOLD / "ray_batch.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    tables = {k[5:]: v for k, v in g.items() if k.startswith("tab__")}
    b = _batcher(tables, cuda_dev)
    _same(b.gather(g["idx"].to(cuda_dev)), {k[5:]: v for k, v in g.items() if k.startswith("out__")})
    b.check()


@pytest.mark.parametrize("feat_dim,feat_h", [(384, 111), (12, 5), (7, 3)])
def test_vs_oracle_reference_shape(cuda_dev, feat_dim, feat_h):
    """DINO-map shape of the reference (111 x 111 x 384), a vector-width and an odd-width case."""
    n_img = 6 if feat_dim == 384 else 4
    tables = RB.synth_tables(n_img, 24, 32, feat_h, feat_dim, seed=3)
    N = tables["all_ray_infos"].shape[0]
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:2048]
    b = _batcher(tables, cuda_dev)
    _same(b.gather(idx.to(cuda_dev)), RB.getitem_batch(tables, idx))
    b.check()


def test_without_features_and_empty(cuda_dev):
    tables = RB.synth_tables(3, 6, 8, 4, 8, seed=9)
    slim = {k: tables[k] for k in ("all_ray_infos", "all_directions", "all_rgbs", "poses")}
    b = _batcher(slim, cuda_dev)
    idx = torch.arange(0, 144, 5)
    got = b.gather(idx.to(cuda_dev))
    assert set(got) == {"ray_infos", "directions", "img_idx", "c2w", "rgbs"}     # datasets/phototourism.py:429
    _same(got, RB.getitem_batch(slim, idx))
    empty = _batcher(tables, cuda_dev).gather(torch.zeros(0, dtype=torch.int64, device=cuda_dev))
    assert empty["feats"].shape == (0, 8) and empty["img_idx"].shape == (0,)


def test_out_of_range_index_is_reported(cuda_dev):
    tables = RB.synth_tables(2, 4, 4, 3, 8, seed=2)
    b = _batcher(tables, cuda_dev)
    b.gather(torch.tensor([0, 5, 31], device=cuda_dev))
    b.check()
    b.gather(torch.tensor([0, 32], device=cuda_dev))            # N = 32: the reference raises IndexError
    with pytest.raises(IndexError):
        b.check()


def test_epoch_is_a_permutation(cuda_dev):
    """shuffle=True semantics: one epoch visits every ray exactly once; batches are what gather() gives."""
    tables = RB.synth_tables(3, 10, 10, 4, 8, seed=4)
    b = _batcher(tables, cuda_dev, batch_size=64, seed=7)
    assert len(b) == 5
    seen, sizes = [], []
    for batch in b:
        sizes.append(batch["rgbs"].shape[0])
        # rgbs rows are unique (continuous random values): recover the ray index by matching
        d = torch.cdist(batch["rgbs"].cpu().double(), tables["all_rgbs"].double())
        seen.append(d.argmin(1))
    assert sizes == [64, 64, 64, 64, 44]
    assert torch.equal(torch.sort(torch.cat(seen)).values, torch.arange(300))
    again = torch.cat([x["img_idx"] for x in b])
    first = torch.cat([x["img_idx"] for x in _batcher(tables, cuda_dev, batch_size=64, seed=7)])
    assert again.shape == first.shape                 # a second epoch reshuffles; a fresh seed-7 batcher repeats epoch 1


def test_batch_feeds_training_step(cuda_dev):
    """The batch dict is exactly what NeRFSystem.training_step consumes (models/nerf_system.py:151-156)."""
    from upnerf_b200.models.nerf_system import NeRFSystem

    tables = RB.synth_tables(4, 16, 16, 9, 384, seed=6)
    b = _batcher(tables, cuda_dev, batch_size=128, seed=1)
    torch.manual_seed(0)
    s = NeRFSystem({"nerf.N_samples": 16, "nerf.N_importance": 16, "max_steps": 1000}, N_images_train=4, device=cuda_dev)
    s.set_progress(0.3)
    batch = next(iter(b))
    loss = s.training_step(batch, 0)
    assert torch.isfinite(loss).all()
