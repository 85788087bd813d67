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


def _reference_param_order(system, which):
    """Parameters in the order the REFERENCE hands them to optimiser `which` (models/nerf_system.py:340-409,
    utils/optim.py:7-17): 0 = the embedding tables (coarse_a, fine_a, coarse_c, fine_c), then nerf_coarse,
    nerf_fine, transient_net; 1 = depth_scale, se3_refine."""
    if which == 1:
        return [system.depth_scale.weight, system.se3_refine.weight]
    out = []
    for tag in ("a", "c"):
        for lvl in ("coarse", "fine"):
            e = system.embeddings.get(f"{lvl}_{tag}")
            if e is not None:
                out += list(e.parameters())
    for m in system.models.values():
        out += list(m.parameters())
    return out


def _load_reference_optimizer(system, which, opt, state):
    """Map a torch.optim.Adam state_dict in the reference's per-tensor layout onto a FlatAdam: exp_avg /
    exp_avg_sq go into the flat moment buffers at each tensor's offset, the per-tensor step counts become
    the per-class counters (tensors of one liveness class share their history), lr comes from param_groups."""
    group = system.group_main if which == 0 else system.group_pose
    params = _reference_param_order(system, which)
    ids = state["param_groups"][0]["params"]
    if len(ids) != len(params):
        raise ValueError(f"optimizer {which}: {len(ids)} tensors in the checkpoint, {len(params)} here")
    flat = opt.param_groups[0]["params"][0]
    st = opt.state[flat]
    st["exp_avg"].zero_()
    st["exp_avg_sq"].zero_()
    key_of = {id(p): k for p, k in zip(group.params, group.keys)}
    steps = {k: 0 for k in opt.class_steps}
    for pid, p in zip(ids, params):
        ts = state["state"].get(pid)
        if ts is None:          # never received a gradient
            continue
        off, n = group.offsets[id(p)]
        if ts["exp_avg"].numel() != n:
            raise ValueError(f"optimizer {which}: tensor {pid} has {ts['exp_avg'].numel()} elements, expected {n}")
        st["exp_avg"][off:off + n].copy_(ts["exp_avg"].reshape(-1))
        st["exp_avg_sq"][off:off + n].copy_(ts["exp_avg_sq"].reshape(-1))
        k = key_of[id(p)]
        steps[k] = max(steps[k], int(ts["step"]))
    opt.class_steps.update(steps)
    g_src, g_dst = state["param_groups"][0], opt.param_groups[0]
    g_dst["lr"] = float(g_src["lr"])
    if "initial_lr" in g_src:
        g_dst["initial_lr"] = float(g_src["initial_lr"])


def load_checkpoint(system, path, strict=True, resume=True):
    """Load a Lightning-format checkpoint (ours or the reference's) into `system`; with `resume`, also the
    step counter, the schedule progress it implies (models/nerf_system.py:222-228) and the optimiser /
    scheduler states -- in this package's flat layout, or mapped from the reference's per-tensor Adam
    state.  If an optimiser state cannot be restored at all, a warning is issued and the learning rate
    is set to where the ExponentialLR schedule stands at the restored step (moments restart from zero)."""
    import warnings

    from ..optim import FlatAdam

    ckpt = torch.load(path, map_location=torch.device("cpu"), weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    system.load_state_dict(sd, strict=strict)
    if resume and "global_step" in ckpt:
        system.global_step = int(ckpt["global_step"])
        if system.hparams.get("pose.optimize", False):
            system.set_progress(system.global_step / (system.hparams["max_steps"] * 2))
        opts, schs = getattr(system, "_optimizers", []), getattr(system, "_schedulers", [])
        ostates, sstates = ckpt.get("optimizer_states") or [], ckpt.get("lr_schedulers") or []
        iters = system.global_step // max(1, len(opts))
        for i, o in enumerate(opts):
            restored = False
            if len(ostates) == len(opts):
                try:
                    if isinstance(o, FlatAdam) and "class_steps" not in ostates[i]:
                        _load_reference_optimizer(system, i, o, ostates[i])
                    else:
                        o.load_state_dict(ostates[i])
                    restored = True
                except (ValueError, KeyError, RuntimeError) as e:
                    warnings.warn(f"load_checkpoint: optimizer {i} state not restored ({e}); Adam moments restart from zero")
            else:
                warnings.warn(f"load_checkpoint: no state for optimizer {i} in the checkpoint; Adam moments restart from zero")
            if i < len(schs):
                if len(sstates) == len(schs):
                    schs[i].load_state_dict(sstates[i])
                else:
                    schs[i].last_epoch = iters
                if not restored or len(sstates) != len(schs):
                    # put the learning rate where the schedule stands: lr0 * gamma ** iterations
                    g = o.param_groups[0]
                    lr0 = g.get("initial_lr", g["lr"])
                    g["lr"] = lr0 * schs[i].gamma ** iters
                schs[i]._last_lr = [g["lr"] for g in o.param_groups]
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
