"""Checkpoint interop with the reference.

The reference trains a LightningModule, so its checkpoints are `{"state_dict": ..., "hyper_parameters":
..., "global_step": ...}` with state_dict keys `nerf_coarse.*`, `nerf_fine.*`, `embedding_{coarse,fine}_{a,c}
.weight`, `transient_net.*`, `se3_refine.weight`, `depth_scale.weight` (models/nerf_system.py:340-409).
`NeRFSystem` here keeps those attribute names, so the files load both ways:

  * `extract_model_state_dict` / `load_ckpt`  -- the reference's helpers (utils/__init__.py:4-26), same
    signatures and prefix semantics (tto.py and eval.py call them per sub-module);
  * `save_checkpoint` / `load_checkpoint`     -- a whole `NeRFSystem` <-> a Lightning-format file.
"""
from __future__ import annotations

import torch


def extract_model_state_dict(ckpt_path, model_name="model", prefixes_to_ignore=[]):
    """utils/__init__.py:4-19: the sub-dict of keys starting with `model_name`, prefix stripped."""
    checkpoint = torch.load(ckpt_path, map_location=torch.device("cpu"), weights_only=False)
    checkpoint_ = {}
    if "state_dict" in checkpoint:      # a pytorch-lightning checkpoint
        checkpoint = checkpoint["state_dict"]
    for k, v in checkpoint.items():
        if not k.startswith(model_name):
            continue
        k = k[len(model_name) + 1:]
        for prefix in prefixes_to_ignore:
            if k.startswith(prefix):
                print("ignore", k)
                break
        else:
            checkpoint_[k] = v
    return checkpoint_


def load_ckpt(model, ckpt_path, model_name="model", prefixes_to_ignore=[]):
    """utils/__init__.py:22-26."""
    model_dict = model.state_dict()
    model_dict.update(extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore))
    model.load_state_dict(model_dict)


def save_checkpoint(system, path):
    """Write `system` (a NeRFSystem) as a Lightning-format checkpoint the reference's tto.py / eval.py read
    (`checkpoint["state_dict"]["se3_refine.weight"]`, `checkpoint["hyper_parameters"]`: eval.py:13-15)."""
    ckpt = {"state_dict": {k: v.detach().cpu().clone() for k, v in system.state_dict().items()},
            "hyper_parameters": dict(system.hparams), "global_step": int(system.global_step),
            "pytorch-lightning_version": "1.9.0"}
    if getattr(system, "_optimizers", None):
        ckpt["optimizer_states"] = [o.state_dict() for o in system._optimizers]
        ckpt["lr_schedulers"] = [s.state_dict() for s in system._schedulers]
    torch.save(ckpt, path)
    return ckpt


def load_checkpoint(system, path, strict=True, resume=True):
    """Load a Lightning-format checkpoint (ours or the reference's) into `system`; with `resume`, also the
    step counter, the schedule progress it implies (models/nerf_system.py:222-228) and, when the file
    has them in this package's flat-buffer layout, the optimiser / scheduler states."""
    ckpt = torch.load(path, map_location=torch.device("cpu"), weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    system.load_state_dict(sd, strict=strict)
    if resume and "global_step" in ckpt:
        system.global_step = int(ckpt["global_step"])
        if system.hparams.get("pose.optimize", False):
            system.set_progress(system.global_step / (system.hparams["max_steps"] * 2))
        for key, objs in (("optimizer_states", getattr(system, "_optimizers", [])),
                          ("lr_schedulers", getattr(system, "_schedulers", []))):
            states = ckpt.get(key) or []
            if len(states) == len(objs):
                for o, st in zip(objs, states):
                    try:
                        o.load_state_dict(st)
                    except (ValueError, KeyError, RuntimeError):
                        # a reference checkpoint: per-tensor optimiser state, not the flat layout
                        pass
    return ckpt


def evaluate_poses(checkpoint, noised_poses, gt_poses):
    """The pose half of eval.py:13-42: compose the learnt se3 refinement with the start poses and report
    the Procrustes-aligned errors.  Returns (mean rotation error in degrees, mean translation error, error dict)."""
    import math

    from . import metric
    from .metric import _compose_pair

    se3 = checkpoint["state_dict"]["se3_refine.weight"] if "state_dict" in checkpoint else checkpoint["se3_refine.weight"]
    dev = noised_poses.device
    se3 = se3.to(dev).float()
    if se3.is_cuda:
        from .camera import lie
        refine = lie.se3_to_SE3(se3)
    else:
        raise RuntimeError("evaluate_poses: the se3 exponential map runs on CUDA tensors only (no CPU fallback)")
    refine_poses = _compose_pair(refine, noised_poses.float())
    err, aligned, gt = metric.pose_metric(refine_poses, gt_poses.to(dev))
    return float(err["R"].mean()) * 180 / math.pi, float(err["t"].mean()), err
