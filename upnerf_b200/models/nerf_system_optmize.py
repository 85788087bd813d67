"""`NeRFSystemOptimize` -- test-time optimisation (tto.py) with the reference's method names
(models/nerf_system_optmize.py:19-332; the module name keeps the reference's spelling).

What the reference does per held-out test image: load a trained checkpoint, freeze nothing
explicitly but put ONLY a fresh `embedding_fine_a` (and, with `pose_optimize`, `se3_refine`) into
the optimisers (:48-64, :253-266), render with `sched_mult=1.0 / sched_phase=2` and both networks'
`encode_candidate=False` (:84-112, :265-266), and minimise `mse(s_rgb_fine, rgbs)` (:128).
`validation_step` renders the whole image in `val.chunk_size` chunks without gradients (:152-169).

Here the same step runs on the CUDA path with the dead work removed -- results are identical
because none of it reaches an optimiser:
  * the networks are frozen: `upnerf_render_bwd` gets `d_params = NULL` and skips every
    weight-gradient launch;
  * the coarse pass receives no gradient at all (the loss reads `s_rgb_fine`, the fine depths come
    from detached coarse weights), so its backward is skipped;
  * without `pose_optimize` nothing upstream of the rgb head's per-ray bias needs a gradient: the
    backward stops after the appearance-embedding gradient (no trunk backward).
Dataset handling, SSIM / LPIPS, image logging and the pose pre-alignment against ground-truth
cameras (:268-332) stay with the caller (out of the hot path).
"""
from __future__ import annotations

from collections import defaultdict

import torch
from torch import nn

from ..optim import FlatAdam
from ..utils import ray as ray_utils
from .nerf_system import FlatGroup, NeRFSystem, allreduce_mean_
from .rendering import render_rays


def extract_model_state_dict(checkpoint, model_name="model", prefixes_to_ignore=()):
    """utils/__init__.py:4-20 on an already loaded checkpoint (a Lightning dict with "state_dict"
    or a plain state dict): the entries of one sub-module, prefix stripped."""
    sd = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    out = {}
    for k, v in sd.items():
        if not k.startswith(model_name):
            continue
        k = k[len(model_name) + 1:]
        if any(k.startswith(p) for p in prefixes_to_ignore):
            continue
        out[k] = v
    return out


class NeRFSystemOptimize(NeRFSystem):
    def __init__(self, hparams, N_images_train=None, N_images_test=None, checkpoint=None, device="cuda"):
        hp = dict(hparams)
        hp.setdefault("pose_optimize", False)
        hp.setdefault("optimize_num", 0)
        super().__init__(hp, None, device)
        self.best_psnr = 0
        if N_images_train is not None:
            self.model_setup(N_images_train, N_images_test, checkpoint)
            self.configure_optimizers()

    # ------------------------------------------------------------------ construction
    @torch.no_grad()
    def model_setup(self, N_images=None, N_images_test=None, checkpoint=None):
        """models/nerf_system_optmize.py:253-266: the training-time modules, the two NeRFs loaded
        from the checkpoint, a FRESH fine appearance table with one row per test image."""
        super().model_setup(N_images)
        hp = self.hparams
        if checkpoint is not None:
            if isinstance(checkpoint, str):
                checkpoint = torch.load(checkpoint, map_location="cpu")
            for name in ("nerf_coarse", "nerf_fine") if self.fine else ("nerf_coarse",):
                net = getattr(self, name)
                own = net.state_dict()
                for k, v in extract_model_state_dict(checkpoint, model_name=name).items():
                    own[k].copy_(v)            # in place: parameters are views of the flat buffer
        n_test = N_images_test if N_images_test is not None else N_images
        self.N_images_test = n_test
        self.embedding_fine_a = nn.Embedding(n_test, hp["nerf.appearance_dim"]).to(self._device)
        self.embeddings["fine_a"] = self.embedding_fine_a
        self.models_to_train = [self.embedding_fine_a]
        self.nerf_coarse.encode_candidate = False
        if self.fine:
            self.nerf_fine.encode_candidate = False
        # nothing of the training-time flat buffer is optimised here: freeze it (the kernels then skip
        # its weight gradients); se3_refine / depth_scale live in group_pose and stay trainable
        self.group_main.flat.requires_grad_(False)
        for p in self.group_main.params:
            p.requires_grad_(False)
        self.depth_scale.weight.requires_grad_(False)
        self.group_tto = FlatGroup([self.embedding_fine_a.weight], self._device, ["always"])
        self._grad_sinks = {"fine_a": self.group_tto.grad_of([self.embedding_fine_a.weight])}

    def configure_optimizers(self):
        """models/nerf_system_optmize.py:48-64 (no schedulers): with `pose_optimize` Adam(5e-3) on the
        appearance table + Adam(1e-4) on se3_refine; otherwise AdamW(1e-1) on the appearance table."""
        hp = self.hparams
        if hp["pose_optimize"]:
            self.optimizer = FlatAdam(self.group_tto.flat, self.group_tto.segments(), lr=5e-3, eps=1e-8)
            self.optimizer_pose = FlatAdam(self.group_pose.flat, self.group_pose.segments(), lr=1e-4, eps=1e-8)
            self.optimizer_pose.set_live({"cand": False, "always": True})     # depth_scale is not optimised
            self._optimizers = [self.optimizer, self.optimizer_pose]
        else:
            self.optimizer = FlatAdam(self.group_tto.flat, self.group_tto.segments(), lr=1e-1, eps=1e-8,
                                      weight_decay=1e-2)                      # torch.optim.AdamW defaults
            self._optimizers = [self.optimizer]
        self._schedulers = []
        return self._optimizers

    # ------------------------------------------------------------------ forward
    def forward(self, rays, img_idx, train=True, rng=None):
        """models/nerf_system_optmize.py:84-112."""
        hp = self.hparams
        B = rays.shape[0]
        chunk = B if train else hp["val.chunk_size"]
        results = defaultdict(list)
        for i in range(0, B, chunk):
            part = render_rays(models=self.models, embeddings=self.embeddings, rays=rays[i:i + chunk],
                               img_idx=img_idx[i:i + chunk], sched_mult=1.0, sched_phase=2,
                               N_samples=hp["nerf.N_samples"], use_disp=hp["nerf.use_disp"],
                               perturb=hp["nerf.perturb"] if train else 0, N_importance=hp["nerf.N_importance"],
                               white_back=self.white_back, encode_feat=hp["nerf.feat_dim"] > 0,
                               validation=not train, precision=hp["kernel.precision"], rng=rng,
                               grad_sink=self._grad_sinks if (train and B == chunk) else None)
            for k, v in part.items():
                results[k].append(v)
        return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in results.items()}

    def _rays(self, batch):
        if self.hparams["pose_optimize"]:
            return ray_utils.refine_rays(self.se3_refine.weight, batch["img_idx"], batch["c2w"],
                                         batch["directions"], batch["ray_infos"])
        o, d = ray_utils.get_rays(batch["directions"], batch["c2w"])
        return torch.cat([o, d, batch["ray_infos"]], 1)

    # ------------------------------------------------------------------ one optimisation step
    def training_step(self, batch, batch_nb=0, rng=None):
        """models/nerf_system_optmize.py:114-150."""
        rgbs = batch["rgbs"]
        rays = self._rays(batch)
        results = self(rays, batch["img_idx"], rng=rng)
        typ = "fine" if self.fine else "coarse"
        loss = ((results[f"s_rgb_{typ}"] - rgbs) ** 2).mean()
        self.group_tto.zero_grad()
        if self.hparams["pose_optimize"]:
            self.group_pose.zero_grad()
        loss.backward()
        allreduce_mean_(self.group_tto.flat.grad)
        if self.hparams["pose_optimize"]:
            allreduce_mean_(self.group_pose.flat.grad)
        for opt in self._optimizers:
            opt.step()
        with torch.no_grad():
            psnr_ = -10 * torch.log10(loss.detach())
        self.log("lr", self.optimizer.param_groups[0]["lr"])
        if self.hparams["pose_optimize"]:
            self.log("lr_pose", self.optimizer_pose.param_groups[0]["lr"])
        self.log("train/loss", loss.detach())
        self.log("train/psnr", psnr_)
        self.global_step += len(self._optimizers)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_nb=0):
        """models/nerf_system_optmize.py:152-203 (render + loss + psnr; SSIM / LPIPS / file output are
        the caller's): `batch` holds ONE image, optionally with the DataLoader's leading 1."""
        b = {k: (v[0] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == 1 and k != "img_wh" else v)
             for k, v in batch.items()}
        rays = self._rays(b)          # one (3,4) pose for the whole image, or one per ray
        results = self(rays, b["img_idx"], train=False)
        typ = "fine" if self.fine else "coarse"
        loss = ((results[f"s_rgb_{typ}"] - b["rgbs"]) ** 2).mean()
        psnr_ = -10 * torch.log10(loss)
        self.log("val/loss", loss)
        self.log("val/psnr", psnr_)
        if float(psnr_) > float(self.best_psnr):
            self.best_psnr = psnr_
        return {"loss": loss, "psnr": psnr_, "s_rgb": results[f"s_rgb_{typ}"], "s_depth": results[f"s_depth_{typ}"]}
