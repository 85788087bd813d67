"""Deterministic synthetic tensors shared by the golden-fixture generator, the tests and
the benchmark (TEST INFRASTRUCTURE; see oracle/upnerf_oracle.py for the import rule).

Weights come from numpy's PCG64 stream, not from torch's initialisers, so that fixtures
stay tiny: a golden file stores only the seed, the inputs and the reference outputs, and
the full-size (W=256) network is regenerated bit-identically wherever the test runs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .upnerf_oracle import NerfConfig


def uniform(shape, seed: int, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    a = rng.random(tuple(shape), dtype=np.float64)
    return torch.from_numpy((lo + (hi - lo) * a).astype(np.float32))


def nerf_param_shapes(cfg: NerfConfig) -> list[tuple[str, tuple]]:
    """Parameter names and shapes in reference state_dict order (models/nerf.py:36-78)."""
    W, in_xyz, in_dir = cfg.W, 6 * cfg.xyz_L + 3, 6 * cfg.dir_L + 3
    out = [("progress", ())]
    for i in range(cfg.D):
        k = in_xyz if i == 0 else (W + in_xyz if i in cfg.skips else W)
        out += [(f"xyz_encoding_{i + 1}.0.weight", (W, k)), (f"xyz_encoding_{i + 1}.0.bias", (W,))]
    out += [("xyz_encoding_final.weight", (W, W)), ("xyz_encoding_final.bias", (W,)),
            ("share_sigma.0.weight", (1, W)), ("share_sigma.0.bias", (1,))]
    if cfg.encode_feat:
        out += [("feat_share_layer.weight", (cfg.feat_dim, W)), ("feat_share_layer.bias", (cfg.feat_dim,))]
        rgb_in = cfg.feat_dim + in_dir
    else:
        rgb_in = W + in_dir
    rgb_in += cfg.appearance_dim
    out += [("rgb_share_layer.0.weight", (W // 2, rgb_in)), ("rgb_share_layer.0.bias", (W // 2,)),
            ("rgb_share_layer.2.weight", (3, W // 2)), ("rgb_share_layer.2.bias", (3,))]
    if cfg.encode_candidate:
        out += [("candidate_encoding.0.weight", (W // 2, W + cfg.candidate_dim)),
                ("candidate_encoding.0.bias", (W // 2,)),
                ("candidate_encoding.2.weight", (W // 2, W // 2)), ("candidate_encoding.2.bias", (W // 2,)),
                ("candidate_sigma.0.weight", (1, W // 2)), ("candidate_sigma.0.bias", (1,))]
        if cfg.encode_feat:
            out += [("feat_candidate_layer.weight", (cfg.feat_dim, W // 2)),
                    ("feat_candidate_layer.bias", (cfg.feat_dim,))]
        else:
            out += [("rgb_candidate_layer.weight", (3, W // 2)), ("rgb_candidate_layer.bias", (3,))]
    return out


def nerf_state(cfg: NerfConfig, seed: int, progress: float = 0.0, gain: float = 1.6) -> dict:
    """Synthetic NeRF state_dict: U(-b, b), b = gain/sqrt(fan_in) (gain>1 keeps activations alive
    through 10 layers so that sigmas/weights are far from degenerate)."""
    sd = {}
    for j, (name, shape) in enumerate(nerf_param_shapes(cfg)):
        if name == "progress":
            sd[name] = torch.tensor(float(progress))
            continue
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:  # bias: look at the matching weight
            fan_in = dict(nerf_param_shapes(cfg))[name.replace(".bias", ".weight")][1]
        b = gain / math.sqrt(fan_in)
        if name.endswith(".bias"):
            b = 0.3 / math.sqrt(fan_in)
        sd[name] = uniform(shape, seed * 1000 + j, -b, b)
    return sd


def embeddings(n_img: int, cfg: NerfConfig, seed: int) -> dict:
    emb = {}
    for j, which in enumerate(("coarse", "fine")):
        if cfg.encode_appearance:
            emb[f"{which}_a"] = uniform((n_img, cfg.appearance_dim), seed * 100 + j, -1.5, 1.5)
        if cfg.encode_candidate:
            emb[f"{which}_c"] = uniform((n_img, cfg.candidate_dim), seed * 100 + 10 + j, -1.5, 1.5)
    return emb


def ray_batch(R: int, n_img: int, seed: int, near: float = 0.1, far: float = 5.0,
              random_pose: bool = True) -> dict:
    """Training-batch dict of SURVEY.md 8(a0) (datasets/phototourism.py:420-454 layout)."""
    g = np.random.default_rng(seed)
    px = g.integers(0, 512, R)
    py = g.integers(0, 384, R)
    directions = np.stack([(px - 256) / 400.0, -(py - 192) / 400.0, -np.ones(R)], -1).astype(np.float32)
    img_idx = torch.from_numpy(g.integers(0, n_img, R).astype(np.int64))
    if random_pose:
        from .upnerf_oracle import se3_exp

        c2w_img = se3_exp(uniform((n_img, 6), seed + 7, -0.25, 0.25))
    else:
        c2w_img = torch.eye(3, 4).expand(n_img, 3, 4).contiguous()
    feats = uniform((R, 384), seed + 3, -1, 1)
    feats = feats / feats.norm(dim=-1, keepdim=True)
    return {
        "ray_infos": torch.tensor([[near, far]], dtype=torch.float32).repeat(R, 1),
        "directions": torch.from_numpy(directions),
        "c2w": c2w_img[img_idx].contiguous(),
        "rgbs": uniform((R, 3), seed + 1, 0, 1),
        "feats": feats,
        "img_idx": img_idx,
        "inv_depths": uniform((R,), seed + 2, 1 / far, 1 / near),
    }


def transient_param_shapes(n_img: int, transient_dim: int = 128, feat_dim: int = 384) -> list[tuple[str, tuple]]:
    """TransientNet parameters in reference state_dict order (models/transient_net.py:6-25)."""
    out = [("embedding_t.weight", (n_img, transient_dim))]
    k = feat_dim
    for i in (0, 2, 4, 6):
        out += [(f"feat_encoder.{i}.weight", (256, k)), (f"feat_encoder.{i}.bias", (256,))]
        k = 256
    out += [("final_encoder.weight", (256, 256)), ("final_encoder.bias", (256,)),
            ("t_encoder.0.weight", (128, 256 + transient_dim)), ("t_encoder.0.bias", (128,)),
            ("alpha_layer.0.weight", (1, 256)), ("alpha_layer.0.bias", (1,)),
            ("beta_layer.0.weight", (1, 128)), ("beta_layer.0.bias", (1,)),
            ("rgb_layer.0.weight", (3, 128)), ("rgb_layer.0.bias", (3,))]
    return out


def transient_state(n_img: int, seed: int) -> dict:
    sd = {}
    for j, (name, shape) in enumerate(transient_param_shapes(n_img)):
        fan = shape[1] if len(shape) == 2 else 64
        b = 1.5 / math.sqrt(fan)
        sd[name] = uniform(shape, seed * 1000 + j, -b, b)
    return sd
