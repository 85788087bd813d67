"""`NeRFSystem` -- the train-step driver with the reference's method names and hparams keys
(models/nerf_system.py:22-461), re-hosted on the CUDA path.

It is a plain `nn.Module` (pytorch-lightning is not a dependency here); `forward`,
`training_step`, `model_setup`, `configure_optimizers` and `get_schedule_mult` keep the
reference's signatures so a Lightning subclass can mix it in unchanged.  Differences that are
deliberate and invisible to results:
  * every trainable tensor is a VIEW into one of two flat fp32 buffers (networks+embeddings /
    pose+depth-scale), with gradients in matching flat buffers -- the CUDA kernels accumulate
    straight into them, one fused Adam updates each, and data-parallel training needs exactly
    one all-reduce per buffer per step (the reference's DDP buckets, train.py:72);
  * `progress` is mirrored on the host, so there is no per-step `.item()` sync
    (models/nerf_system.py:180);
  * pose refinement + ray casting is one fused kernel (`refine_rays`).
"""
from __future__ import annotations

import math
import os
from collections import defaultdict

import torch
import torch.distributed as dist
from torch import nn

from .. import _lib as _L
from ..losses import UPNeRFLoss, fused_tail
from ..optim import FlatAdam
from ..utils import ray as ray_utils
from .nerf import NeRF
from .rendering import render_rays
from .transient_net import TransientNet


def default_hparams() -> dict:
    """The reference's shipped defaults (configs/default.yaml + configs/brandenburg_gate.yaml), flat
    dotted keys like configs/config.py produces."""
    return {
        "seed": 42, "num_gpus": 1, "debug": False,
        "nerf.N_samples": 128, "nerf.N_importance": 128, "nerf.N_emb_xyz": 10, "nerf.N_emb_dir": 4,
        "nerf.near": 0.1, "nerf.far": 5.0, "nerf.appearance_dim": 48, "nerf.candidate_dim": 16,
        "nerf.feat_dim": 384, "nerf.use_disp": False, "nerf.perturb": 1.0,
        "t_net.beta_min": 0.1, "t_net.transient_dim": 128, "t_net.feat_dim": 384,
        "loss.depth_mult": 1e-3, "loss.alpha_reg": 1.0,
        "optimizer.type": "adam", "optimizer.lr": 5e-4, "optimizer.scheduler.type": "ExponentialLR",
        "optimizer.scheduler.lr_end": 5e-5,
        "optimizer_pose.type": "adam", "optimizer_pose.lr": 2e-3,
        "optimizer_pose.scheduler.type": "ExponentialLR", "optimizer_pose.scheduler.lr_end": 1e-5,
        "max_steps": 600000, "train.batch_size": 2048, "train.log_pose_interval": 3000,
        "val.chunk_size": 4096,
        "pose.optimize": True, "pose.c2f": (0.1, 0.5), "pose.noise": -1,
        "candidate_schedule": (0.1, 0.5),
        "kernel.precision": "bf16",
        # TransientNet on a second CUDA stream beside the render (UPNERF_TNET_SIDE=0 turns it off)
        "kernel.side_stream": os.environ.get("UPNERF_TNET_SIDE", "1") != "0",
        "kernel.fused_tail": True,      # depth correction + loss + its backward + psnr as one launch (CUDA only)
        # the whole optimisation step as ONE CUDA-graph replay (UPNERF_CUDA_GRAPH=0 turns it off); the first
        # `kernel.cuda_graph_warmup` steps run eagerly (lazy initialisation must not happen under capture)
        "kernel.cuda_graph": os.environ.get("UPNERF_CUDA_GRAPH", "1") != "0",
        "kernel.cuda_graph_warmup": 3,
        # data-parallel runs capture the two NCCL all-reduces of the step into the graph as well
        "kernel.cuda_graph_ddp": os.environ.get("UPNERF_CUDA_GRAPH_DDP", "1") != "0",
    }


class FlatGroup:
    """Parameters re-homed as views of one flat fp32 buffer, gradients likewise."""

    def __init__(self, params, device, keys=None, grad_storage=None):
        """`grad_storage`: an fp32 buffer of the right size to use as the flat gradient (so that several
        groups can share ONE allocation and be all-reduced with one collective)."""
        self.params = list(params)
        self.keys = list(keys) if keys is not None else ["always"] * len(self.params)
        n = sum(p.numel() for p in self.params)
        self.flat = nn.Parameter(torch.empty(n, device=device, dtype=torch.float32))
        if grad_storage is None:
            grad_storage = torch.zeros(n, device=device, dtype=torch.float32)
        assert grad_storage.numel() == n and grad_storage.dtype == torch.float32 and grad_storage.is_contiguous()
        self.flat.grad = grad_storage
        off = 0
        self.offsets = {}
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat.data[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat.data[off:off + k].view(p.shape)
                p.grad = self.flat.grad[off:off + k].view(p.shape)
                self.offsets[id(p)] = (off, k)
                off += k

    def zero_grad(self):
        self.flat.grad.zero_()

    def check(self):
        """Every parameter (and its .grad) must still be the view of the flat buffers it was made: a later
        `module.to()`, `.float()`, `zero_grad(set_to_none=True)` or `load_state_dict(assign=True)` on a
        sub-module silently re-homes tensors, and the fused optimiser would then update stale memory."""
        es = self.flat.element_size()
        base, gbase = self.flat.data_ptr(), self.flat.grad.data_ptr()
        for p in self.params:
            off, _ = self.offsets[id(p)]
            if p.data_ptr() != base + off * es or p.grad is None or p.grad.data_ptr() != gbase + off * es:
                raise RuntimeError("a parameter of NeRFSystem no longer aliases its flat buffer (was a sub-module "
                                   "moved, cast, or its gradients set to None?); rebuild the system instead")

    def segments(self):
        """[(end_offset, class_key)] runs covering the buffer (the FlatAdam segment table)."""
        return [(self.offsets[id(p)][0] + self.offsets[id(p)][1], k) for p, k in zip(self.params, self.keys)]

    def grad_of(self, params):
        """The slice of the flat GRADIENT buffer covering a run of consecutive parameters."""
        params = list(params)
        a = self.offsets[id(params[0])][0]
        o, k = self.offsets[id(params[-1])]
        return self.flat.grad[a:o + k]

    def slice_of(self, params):
        """Getter of the autograd-visible slice of the flat leaf covering a run of consecutive parameters."""
        params = list(params)
        a = self.offsets[id(params[0])][0]
        o, k = self.offsets[id(params[-1])]
        return lambda: self.flat[a:o + k]


def adam_class(name: str) -> str:
    """Which schedule phases put a parameter into the loss graph (= give it a non-None .grad in the
    reference, models/nerf.py:80-124, models/nerf_system.py:128-144, losses.py:21-64):
    "rgb" = sched_mult > 0, "cand" = sched_mult < 1, "never" (progress; TransientNet.rgb_layer only
    feeds rgb_fine, which no loss term reads), "always"."""
    head = name.split(".")[0]
    if name.endswith("progress"):
        return "never"
    if head == "transient_net":
        return "never" if ".rgb_layer." in name else "rgb"
    if head.startswith("embedding_"):
        return "rgb" if head.endswith("_a") else "cand"
    if head == "depth_scale":
        return "cand"
    if ".rgb_share_layer." in name:
        return "rgb"
    if ".candidate_" in name or ".feat_candidate_layer." in name or ".rgb_candidate_layer." in name:
        return "cand"
    return "always"


def allreduce_mean_(t: torch.Tensor) -> None:
    """DDP gradient semantics (train.py:72): mean over ranks, one collective per flat buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(dist.get_world_size())


class NeRFSystem(nn.Module):
    def __init__(self, hparams, N_images_train=None, device="cuda"):
        super().__init__()
        self.hparams = dict(default_hparams(), **dict(hparams))
        self.candidate_schedule = self.hparams["candidate_schedule"]
        self.fine = self.hparams["nerf.N_importance"] > 0
        self.loss = UPNeRFLoss(depth_mult=self.hparams["loss.depth_mult"], alpha_reg=self.hparams["loss.alpha_reg"],
                               encode_feat=self.hparams["nerf.feat_dim"] > 0, fine=self.fine)
        self.global_step = 0
        self.logged = {}
        self._device = torch.device(device)
        self._progress = 0.0
        self._tail_ws = None
        self._side_stream = None
        self._steps_seen = 0
        self._graphs = {}               # key -> _StepGraph (captured optimisation steps)
        self._graph_pool = None
        self.graph_replays = 0
        if self.hparams["nerf.feat_dim"] <= 0 and self.hparams["nerf.candidate_dim"] > 0:
            raise NotImplementedError(
                "upnerf_b200: encode_feat=False (nerf.feat_dim <= 0) together with a candidate head -- the "
                "reference's c_rgb_* / l_c_rgb_* path (models/rendering.py:134-150) -- is not implemented; "
                "use nerf.feat_dim > 0, or nerf.candidate_dim = 0 (INTEGRATION.md section 3)")
        self.white_back = False
        if N_images_train is not None:
            self.model_setup(N_images_train)
            self.configure_optimizers()

    # ------------------------------------------------------------------ construction
    def model_setup(self, N_images=None):
        """models/nerf_system.py:340-409 (same attribute names and state_dict keys)."""
        hp = self.hparams
        if N_images is None:
            N_images = self.train_dataset.N_images_train
        self.N_images = N_images
        self.embeddings = {}
        for which in ("coarse", "fine") if self.fine else ("coarse",):
            for tag, dim in (("a", hp["nerf.appearance_dim"]), ("c", hp["nerf.candidate_dim"])):
                if dim > 0:
                    e = nn.Embedding(N_images, dim)
                    setattr(self, f"embedding_{which}_{tag}", e)
                    self.embeddings[f"{which}_{tag}"] = e
        mk = lambda typ: NeRF(typ, encode_feat=hp["nerf.feat_dim"] > 0, feat_dim=hp["nerf.feat_dim"],
                              xyz_L=hp["nerf.N_emb_xyz"], dir_L=hp["nerf.N_emb_dir"],
                              appearance_dim=hp["nerf.appearance_dim"], candidate_dim=hp["nerf.candidate_dim"],
                              c2f=hp["pose.c2f"])
        self.nerf_coarse = mk("coarse")
        self.models = {"nerf_coarse": self.nerf_coarse}
        if self.fine:
            self.nerf_fine = mk("fine")
            self.models["nerf_fine"] = self.nerf_fine
        self.transient_net = TransientNet(N_images=N_images, beta_min=hp["t_net.beta_min"],
                                          trasient_dim=hp["t_net.transient_dim"], feat_dim=hp["t_net.feat_dim"])
        self.models["transient_network"] = self.transient_net
        # bf16 mode: TransientNet forward/backward on the tensor-core path, gradients accumulated straight into
        # the flat gradient buffer (its parameters' .grad are views of it)
        self.transient_net.precision = "bf16" if hp["kernel.precision"] in ("bf16", "bfloat16") else "fp32"
        self.transient_net.grad_sink = True
        self.se3_refine = nn.Embedding(N_images, 6)
        nn.init.zeros_(self.se3_refine.weight)
        self.depth_scale = nn.Embedding(N_images, 2)
        nn.init.zeros_(self.depth_scale.weight)
        self.group_main = None
        self.to(self._device)
        # flat buffers: group 0 = reference optimizer 0 (networks + embeddings), group 1 = pose
        # Layout [nerf_fine | nerf_coarse | transient_net | embeddings] + [depth_scale | se3_refine], gradients of
        # both groups in ONE allocation: the fine network's slice is final first in backward (its all-reduce
        # then runs under the coarse pass), everything else goes out as one more collective.
        main, keys = [], []
        order = (["nerf_fine"] if self.fine else []) + ["nerf_coarse", "transient_network"]
        for mname in order:
            m = self.models[mname]
            attr = {"nerf_coarse": "nerf_coarse", "nerf_fine": "nerf_fine", "transient_network": "transient_net"}[mname]
            for pname, p in m.named_parameters():
                main.append(p)
                keys.append(adam_class(f"{attr}.{pname}"))
        for ename, e in self.embeddings.items():
            main += list(e.parameters())
            keys.append(adam_class(f"embedding_{ename}.weight"))
        n_main = sum(p.numel() for p in main)
        n_pose = self.depth_scale.weight.numel() + self.se3_refine.weight.numel()
        n_pad = (n_main + 63) // 64 * 64          # the pose group starts 256-byte aligned (fused Adam: 16-byte vectors)
        self._grad_all = torch.zeros(n_pad + n_pose, device=self._device, dtype=torch.float32)
        self.group_main = FlatGroup(main, self._device, keys, grad_storage=self._grad_all[:n_main])
        self.group_pose = FlatGroup([self.depth_scale.weight, self.se3_refine.weight], self._device,
                                    [adam_class("depth_scale.weight"), adam_class("se3_refine.weight")],
                                    grad_storage=self._grad_all[n_pad:])
        self._n_fine = sum(p.numel() for p in self.nerf_fine.parameters()) if self.fine else 0
        self._pending_reduce = None
        # the render backward accumulates straight into these slices of the flat gradient buffer
        self._grad_sinks = {}
        for which, m in (("coarse", self.nerf_coarse), *((("fine", self.nerf_fine),) if self.fine else ())):
            m._upnerf_flat = self.group_main.slice_of(m.parameters())
            self._grad_sinks[which] = self.group_main.grad_of(m.parameters())
        for ename, e in self.embeddings.items():
            self._grad_sinks[ename] = self.group_main.grad_of([e.weight])

    def _apply(self, fn, recurse=True):
        # nn.Module.to()/.float()/.cuda() replace parameter storage; once the flat buffers exist that would
        # detach every view from them (FlatAdam and the gradient sinks would update stale memory)
        if getattr(self, "group_main", None) is not None:
            raise RuntimeError("NeRFSystem: parameters are views of flat buffers; construct it with device=... "
                               "instead of moving or casting it afterwards")
        return super()._apply(fn, recurse)

    def zero_grad(self, set_to_none=False):
        """Gradients live in the flat buffers: they are zeroed in place, never set to None."""
        self._grad_all.zero_()

    # ------------------------------------------------------------------ data-parallel gradient mean
    def _ddp_active(self):
        return (not self.hparams.get("kernel.skip_allreduce", False) and dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1)

    def _reduce_fine_async(self):
        """Called between the fine and the coarse backward (`render_rays(after_fine_bwd=...)`): the fine
        network's gradient slice is final, its mean over ranks starts now on NCCL's stream."""
        if self._n_fine and self._grad_all.is_cuda:
            self._pending_reduce = dist.all_reduce(self._grad_all[:self._n_fine], op=dist.ReduceOp.AVG, async_op=True)

    def _reduce_gradients(self):
        """DDP semantics (train.py:72): mean over ranks of every gradient -- the rest of the single gradient
        buffer in one collective, then wait for the fine slice started earlier."""
        if not self._ddp_active():
            return
        lo = self._n_fine if self._pending_reduce is not None else 0
        if not self.hparams["pose.optimize"]:
            allreduce_mean_(self._grad_all[lo:self.group_main.flat.numel()])
        else:
            allreduce_mean_(self._grad_all[lo:])
        if self._pending_reduce is not None:
            self._pending_reduce.wait()
            self._pending_reduce = None

    def load_state_dict(self, sd, strict=True):
        # parameters are views of the flat buffers: copy in place so the views stay valid
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state_dict mismatch: missing {missing[:5]} unexpected {unexpected[:5]}")
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v)
        self._progress = float(self.nerf_coarse.progress.data.item())

    def configure_optimizers(self):
        """Two Adam(eps=1e-8) + ExponentialLR pairs (models/nerf_system.py:41-73, utils/optim.py:20-44)."""
        hp = self.hparams
        fused = self._device.type == "cuda"

        def make(prefix, group):
            if hp[f"{prefix}.type"] != "adam":
                raise NotImplementedError(hp[f"{prefix}.type"])
            if fused:      # one upnerf_adam_step launch, per-segment liveness like the per-tensor reference
                opt = FlatAdam(group.flat, group.segments(), lr=hp[f"{prefix}.lr"], eps=1e-8)
            else:
                opt = torch.optim.Adam([group.flat], lr=hp[f"{prefix}.lr"], eps=1e-8)
            gamma = (hp[f"{prefix}.scheduler.lr_end"] / hp[f"{prefix}.lr"]) ** (1.0 / hp["max_steps"])
            return opt, torch.optim.lr_scheduler.ExponentialLR(opt, gamma=gamma)

        self.optimizer, self.scheduler = make("optimizer", self.group_main)
        self._optimizers, self._schedulers = [self.optimizer], [self.scheduler]
        if hp["pose.optimize"]:
            self.optimizer_pose, self.scheduler_pose = make("optimizer_pose", self.group_pose)
            self._optimizers.append(self.optimizer_pose)
            self._schedulers.append(self.scheduler_pose)
        return self._optimizers, self._schedulers

    def optimizers(self):
        return self._optimizers

    def lr_schedulers(self):
        return self._schedulers

    def log(self, key, value, **kw):
        self.logged[key] = value

    # ------------------------------------------------------------------ schedule
    def get_schedule_mult(self, progress):
        """models/nerf_system.py:452-461."""
        s, e = self.candidate_schedule
        if progress < s:
            return 0
        if progress > e:
            return 1
        return (1 - math.cos(math.pi * (progress - s) / (e - s))) / 2

    def _live_terms(self, m):
        """Keys UPNeRFLoss returns in this phase (losses.py:21-64)."""
        levels = ("c", "f") if self.fine else ("c",)
        keys = []
        for tag in levels:
            if m < 1:
                keys += [f"l_depth_{tag}", f"l_feat_{tag}"]
            if m > 0:
                keys += [f"l_rgb_{tag}"] + (["l_beta", "l_alpha"] if tag == "f" else [])
        return keys

    def set_progress(self, progress: float, fill=True):
        # the reference keeps progress in an fp32 Parameter and reads it back with .item()
        # (models/nerf_system.py:180,220-226): the schedule sees the fp32-rounded value
        # (`fill=False`: the graphed step has already written the device copies)
        self._progress = float(torch.tensor(float(progress), dtype=torch.float32))
        if not fill:
            return
        self.nerf_coarse.progress.data.fill_(self._progress)
        if self.fine:
            self.nerf_fine.progress.data.fill_(self._progress)

    # ------------------------------------------------------------------ forward
    def forward(self, rays, feats, img_idx, sched_mult, train=True, rng=None, blend=True):
        """Chunked render + transient blend (models/nerf_system.py:93-148).  `blend=False` (the fused
        training path) skips `rgb_coarse` / `rgb_fine`, which no loss term reads (losses.py:39-64)."""
        hp = self.hparams
        B = rays.shape[0]
        chunk = B if train else hp["val.chunk_size"]
        # TransientNet reads only (feats, img_idx): on the CUDA path it runs on a second stream beside the
        # render (autograd runs its backward on that stream too, beside the render backward)
        t, side = None, None
        if sched_mult > 0 and rays.is_cuda and hp["kernel.side_stream"]:
            main = torch.cuda.current_stream(rays.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=rays.device)
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                t = self.transient_net(feats, img_idx)
            feats.record_stream(side)
            img_idx.record_stream(side)
        kw = dict(models=self.models, embeddings=self.embeddings, sched_mult=sched_mult,
                  sched_phase=0 if sched_mult == 0 else (2 if sched_mult == 1 else 1),
                  N_samples=hp["nerf.N_samples"], use_disp=hp["nerf.use_disp"],
                  perturb=hp["nerf.perturb"] if train else 0, N_importance=hp["nerf.N_importance"],
                  white_back=self.white_back, encode_feat=hp["nerf.feat_dim"] > 0,
                  validation=not train, precision=hp["kernel.precision"])
        if (not train and hp["kernel.cuda_graph"] and rays.is_cuda and not torch.is_grad_enabled() and rng is None
                and B >= 2 * chunk):
            # inference (config 5: a 1920x1080 image is 507 chunks of val.chunk_size rays): one captured chunk,
            # replayed per chunk -- the eager per-chunk host work (30 launches through Python) took longer than
            # the chunk's 1.2 ms on the device
            results = self._render_chunks_graphed(rays, img_idx, chunk, kw)
        else:
            results = defaultdict(list)
            for i in range(0, B, chunk):
                part = render_rays(rays=rays[i:i + chunk], img_idx=img_idx[i:i + chunk], rng=rng,
                                   grad_sink=self._grad_sinks if (train and B == chunk) else None,
                                   after_fine_bwd=self._reduce_fine_async if (train and B == chunk and self._ddp_active())
                                   else None, **kw)
                for k, v in part.items():
                    results[k].append(v)
            results = {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in results.items()}
        if sched_mult > 0:
            if t is None:
                t = self.transient_net(feats, img_idx)
            else:
                main.wait_stream(side)
                for v in t.values():
                    v.record_stream(main)
            a, c = t["alpha"], t["rgb"]
            if blend:
                results["rgb_coarse"] = results["s_rgb_coarse"] * (1 - a.detach()) + c.detach() * a.detach()
                if self.fine:
                    results["rgb_fine"] = results["s_rgb_fine"] * (1 - a) + c * a
            results["t_beta"], results["t_alpha"] = t["beta"], a
        return results

    # ------------------------------------------------------------------ one optimisation step
    def _check_batch(self, img_idx):
        """Cheap invariants, checked on the first step and then rarely (one host sync): indices in range
        (nn.Embedding would raise; the kernels index raw tables) and the flat-buffer aliasing."""
        if img_idx.numel() and not (0 <= int(img_idx.min()) and int(img_idx.max()) < self.N_images):
            raise IndexError(f"img_idx out of range [0, {self.N_images})")
        self.group_main.check()
        self.group_pose.check()

    def _live_classes(self, sched_mult):
        return {"always": True, "never": False, "rgb": sched_mult > 0 or self.hparams["nerf.feat_dim"] <= 0,
                "cand": sched_mult < 1}

    def training_step(self, batch, batch_nb=0, rng=None):
        """models/nerf_system.py:150-229."""
        hp = self.hparams
        img_idx = batch["img_idx"]
        if img_idx.dtype != torch.int64 or not img_idx.is_contiguous():
            img_idx = img_idx.contiguous().long()       # the kernels read int64 indices
            batch = dict(batch, img_idx=img_idx)
        if self._steps_seen % 256 == 0:
            self._check_batch(img_idx)
        self._steps_seen += 1
        sched_mult = self.get_schedule_mult(self._progress)
        if (hp["kernel.cuda_graph"] and hp["kernel.fused_tail"] and rng is None and img_idx.is_cuda
                and self._steps_seen > hp["kernel.cuda_graph_warmup"] and torch.is_grad_enabled()
                and (not self._ddp_active() or hp["kernel.cuda_graph_ddp"])):
            return self._training_step_graphed(batch, sched_mult)
        loss, loss_d, psnr_ = self._step_body(batch, sched_mult, rng=rng)
        for opt in self._optimizers:
            if isinstance(opt, FlatAdam):
                opt.set_live(self._live_classes(sched_mult))
            opt.step()
        return self._finish_step(loss, loss_d, psnr_)

    def _finish_step(self, loss, loss_d, psnr_):
        """Host bookkeeping of a step: learning-rate schedules, logging, step counters, progress."""
        hp = self.hparams
        for opt, sch in zip(self._optimizers, self._schedulers):
            opt._opt_called = True          # (the graphed step launches the update itself, not through opt.step)
            sch.step()
        self.log("train/loss", loss.detach())
        for k, v in loss_d.items():
            self.log(f"train/{k}", v.detach())
        self.log("train/psnr", psnr_)
        self.global_step += len(self._optimizers)
        if hp["pose.optimize"]:
            self.set_progress(self.global_step / (hp["max_steps"] * 2), fill=not self._in_graph_step)
        return loss

    _in_graph_step = False

    def _step_body(self, batch, sched_mult, rng=None, scal=None):
        """Pose refinement -> render -> TransientNet -> loss -> backward -> gradient mean: everything of a step
        that runs on the device before the optimiser update.  `scal` (graph capture): device tensor holding the
        per-step scalars, [0] = sched_mult."""
        hp = self.hparams
        img_idx = batch["img_idx"]
        if hp["pose.optimize"]:
            rays = ray_utils.refine_rays(self.se3_refine.weight, img_idx, batch["c2w"], batch["directions"],
                                         batch["ray_infos"])
        else:
            o, d = ray_utils.get_rays(batch["directions"], batch["c2w"])
            rays = torch.cat([o, d, batch["ray_infos"]], 1)
        if hp["kernel.fused_tail"] and rays.is_cuda:
            # render + TransientNet, then ONE launch for the depth correction (:169-177), UPNeRFLoss
            # (losses.py:21-64), its backward and psnr (:202-207); autograd continues from the
            # gradients that kernel wrote
            results = self(rays, batch["feats"], img_idx, sched_mult, rng=rng, blend=False)
            self._grad_all.zero_()
            if self._tail_ws is None:
                from .. import _lib as L
                self._tail_ws = torch.zeros(L.tail_workspace_bytes(), device=rays.device, dtype=torch.uint8)
            losses, roots, grads = fused_tail(results, batch, self.depth_scale.weight, sched_mult,
                                              depth_mult=hp["loss.depth_mult"], alpha_reg=hp["loss.alpha_reg"],
                                              near=hp["nerf.near"], far=hp["nerf.far"], fine=self.fine,
                                              workspace=self._tail_ws,
                                              sched_mult_dev=None if scal is None else scal[0:1])
            torch.autograd.backward(roots, grads)
            from .._lib import TAIL_TERMS
            live = self._live_terms(sched_mult)
            loss_d = {k: losses[i] for i, k in enumerate(TAIL_TERMS) if k in live}
            loss, psnr_ = losses[8], losses[9]
        else:
            # monocular-depth affine correction (:169-177)
            scale, shift = torch.unbind(self.depth_scale(img_idx), 1)
            inv = batch["inv_depths"] * torch.exp(scale) + shift
            inv = torch.where(inv < 1 / hp["nerf.far"], torch.full_like(inv, 1 / hp["nerf.far"]), inv)
            depth = 1.0 / inv
            depth = torch.where(depth < hp["nerf.near"], torch.full_like(depth, hp["nerf.near"]), depth)
            results = self(rays, batch["feats"], img_idx, sched_mult, rng=rng)
            loss_d = self.loss(results, batch["rgbs"], batch["feats"], depth, sched_mult)
            loss = sum(loss_d.values())
            self._grad_all.zero_()
            loss.backward()
            with torch.no_grad():
                typ = "fine" if self.fine else "coarse"
                if f"s_rgb_{typ}" in results:
                    psnr_ = -10 * torch.log10(((results[f"s_rgb_{typ}"] - batch["rgbs"]) ** 2).mean())
                else:
                    psnr_ = torch.zeros(1)
        self._reduce_gradients()      # (kernel.skip_allreduce: bench.py measures the step without it)
        return loss, loss_d, psnr_

    # ------------------------------------------------------------------ the step as one CUDA-graph replay
    # (reference loop models/nerf_system.py:150-229.)  Everything the host decides per step
    # and that changes continuously -- sched_mult in the loss weights, `progress`, both optimisers' step sizes and
    # bias corrections -- lives in a small device array the host refreshes with one small copy before each replay
    # (from pageable memory: staged at call time, so the host may run ahead of the device); everything discrete (schedule phase, the static/candidate split of the fine samples, live parameter
    # classes, batch shapes) is the key under which a graph is captured.
    def _graph_key(self, batch, sched_mult):
        hp = self.hparams
        phase = 0 if sched_mult == 0 else (2 if sched_mult == 1 else 1)
        two = self.fine and hp["nerf.candidate_dim"] > 0 and 0 < sched_mult < 1
        n_static = round(sched_mult * hp["nerf.N_importance"]) if two else 0
        flags = tuple(tuple(o.live_flags()) for o in self._optimizers if isinstance(o, FlatAdam))
        shapes = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.items()) if torch.is_tensor(v))
        return (phase, n_static, flags, shapes, self._ddp_active(), bool(hp.get("kernel.skip_allreduce", False)))

    def _training_step_graphed(self, batch, sched_mult):
        hp = self.hparams
        for opt in self._optimizers:
            opt.set_live(self._live_classes(sched_mult))
        key = self._graph_key(batch, sched_mult)
        g = self._graphs.get(key)
        if g is None:
            g = self._capture_step(batch, sched_mult)
            if len(self._graphs) >= 4:            # keep a few phases' graphs; they share one memory pool
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        # host half of the step: the scalars of THIS step
        n_sc = FlatAdam.N_SCALARS
        h = [0.0] * (2 + n_sc * len(self._optimizers))
        h[0] = sched_mult
        next_progress = self._progress
        if hp["pose.optimize"]:
            gs = self.global_step + len(self._optimizers)
            next_progress = float(torch.tensor(gs / (hp["max_steps"] * 2), dtype=torch.float32))
        h[1] = next_progress
        for j, opt in enumerate(self._optimizers):
            _, vals = opt.advance()
            h[2 + j * n_sc: 2 + (j + 1) * n_sc] = vals
        g["scal"].copy_(torch.tensor(h, dtype=torch.float32), non_blocking=True)
        # this step's batch -> the graph's static inputs: one fused multi-tensor copy per dtype instead of a launch
        # per tensor (eight 2 us copies in a row were ~50 us of every step)
        groups = {}
        for k, dst in g["batch"].items():
            src = batch[k]
            if src.data_ptr() != dst.data_ptr():
                if src.device == dst.device and src.dtype == dst.dtype and src.shape == dst.shape and src.is_contiguous():
                    d_, s_ = groups.setdefault(dst.dtype, ([], []))
                    d_.append(dst)
                    s_.append(src)
                else:
                    dst.copy_(src, non_blocking=True)
        for d_, s_ in groups.values():
            if len(d_) > 1:
                torch._foreach_copy_(d_, s_, non_blocking=True)
            else:
                d_[0].copy_(s_[0], non_blocking=True)
        g["graph"].replay()
        self.graph_replays += 1
        _L.launch_count_add(g["launches"])
        out = g["losses"].clone()             # the static output is overwritten by the next replay
        loss_d = {k: out[i] for k, i in g["loss_idx"].items()}
        self._in_graph_step = True
        try:
            return self._finish_step(out[8], loss_d, out[9])
        finally:
            self._in_graph_step = False

    def _render_chunks_graphed(self, rays, img_idx, chunk, kw):
        """Chunked no-grad render (models/nerf_system.py:104-126) with the full chunks replayed from ONE captured
        graph: static input buffers, outputs copied into preallocated full-size tensors.  The captured kernels
        read parameters, `progress` and embeddings from their live buffers, so training may continue between
        renders; everything baked into the graph is part of its key."""
        B = rays.shape[0]
        key = ("render", chunk, float(kw["sched_mult"]), kw["N_samples"], kw["N_importance"], bool(kw["use_disp"]),
               kw["precision"], rays.dtype, img_idx.dtype)
        g = self._graphs.get(key)
        if g is None:
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            s_rays = rays[:chunk].contiguous().clone()
            s_idx = img_idx[:chunk].contiguous().clone()
            # every chunk runs in ONE workspace (allocated by this eager call, outside the captures): the first chunk
            # of a render packs the GEMM operands / folded head matrices / c2f weights, the others reuse them
            wsc = {}
            render_rays(rays=s_rays, img_idx=s_idx, workspace_cache=wsc, **kw)   # lazy initialisation outside the capture
            torch.cuda.synchronize(rays.device)
            g = {"rays": s_rays, "idx": s_idx, "ws": wsc}
            for name, reuse in (("first", False), ("rest", True)):
                graph = torch.cuda.CUDAGraph()
                launches0 = _L.launch_count()
                with torch.cuda.graph(graph, pool=self._graph_pool):
                    part = render_rays(rays=s_rays, img_idx=s_idx, workspace_cache=wsc, reuse_packed=reuse, **kw)
                g[name] = {"graph": graph, "part": part, "launches": _L.launch_count() - launches0}
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        first = g["first"]["part"]
        out = {k: torch.empty((B,) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype) for k, v in first.items()}
        n_full = B // chunk
        for c in range(n_full):
            i = c * chunk
            gc = g["first" if c == 0 else "rest"]      # parameters may have changed since the previous render
            g["rays"].copy_(rays[i:i + chunk], non_blocking=True)
            g["idx"].copy_(img_idx[i:i + chunk], non_blocking=True)
            gc["graph"].replay()
            _L.launch_count_add(gc["launches"])
            self.graph_replays += 1
            # the chunk's outputs -> their rows of the full-size tensors: one fused multi-tensor copy
            torch._foreach_copy_([out[k][i:i + chunk] for k in gc["part"]], list(gc["part"].values()), non_blocking=True)
        if n_full * chunk < B:                                      # the ragged last chunk runs eagerly
            i = n_full * chunk
            part = render_rays(rays=rays[i:], img_idx=img_idx[i:], **kw)
            for k, v in part.items():
                out[k][i:].copy_(v)
        return out

    def release_graphs(self):
        """Drop every captured step.  Data-parallel runs MUST call this before
        `torch.distributed.destroy_process_group()`: a captured graph holds NCCL's persistent plans, and
        NCCL's communicator teardown waits for the graphs that reference it."""
        if self._graphs:
            torch.cuda.synchronize(self._device)
        self._graphs.clear()
        self._graph_pool = None

    def _capture_step(self, batch, sched_mult):
        dev = self._device
        n_sc = FlatAdam.N_SCALARS
        n_opt = len(self._optimizers)
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        static = {k: torch.empty_like(v) for k, v in batch.items() if torch.is_tensor(v)}
        for k, v in static.items():
            v.copy_(batch[k])
        scal = torch.zeros(2 + n_opt * n_sc, device=dev, dtype=torch.float32)
        flags = [opt.live_flags() for opt in self._optimizers]
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        launches0 = _L.launch_count()
        with torch.cuda.graph(graph, pool=self._graph_pool):
            loss, loss_d, psnr_ = self._step_body(static, sched_mult, scal=scal)
            losses = loss._base if loss._base is not None else loss
            for j, opt in enumerate(self._optimizers):
                opt.launch(flags[j], dev_scalars=scal[2 + j * n_sc: 2 + (j + 1) * n_sc])
            if self.hparams["pose.optimize"]:
                self.nerf_coarse.progress.data.copy_(scal[1:2].view_as(self.nerf_coarse.progress.data))
                if self.fine:
                    self.nerf_fine.progress.data.copy_(scal[1:2].view_as(self.nerf_fine.progress.data))
        from .._lib import TAIL_TERMS
        live = self._live_terms(sched_mult)
        return {"graph": graph, "batch": static, "scal": scal, "losses": losses,
                "launches": _L.launch_count() - launches0,
                "loss_idx": {k: i for i, k in enumerate(TAIL_TERMS) if k in live}}
