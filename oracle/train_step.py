"""Oracle restatement of one full optimisation step (TEST INFRASTRUCTURE; see upnerf_oracle.py
for the import rule): the reference's NeRFSystem.training_step (models/nerf_system.py:150-229) and
configure_optimizers (:41-73, utils/optim.py:20-44) on the oracle's functions, CPU only."""
from __future__ import annotations

import torch

from . import synth
from . import upnerf_oracle as O

KW = dict(W=256, feat_dim=384, appearance_dim=48, candidate_dim=16, xyz_L=10, dir_L=4, c2f=(0.1, 0.5))


class OracleSystem:
    """The reference training_step (models/nerf_system.py:150-229) restated on oracle functions."""

    def __init__(self, cfgs, sd, n_img, S, NI, max_steps):
        self.cfgs, self.S, self.NI, self.max_steps = cfgs, S, NI, max_steps
        self.p = {k: v.clone().requires_grad_(k.split(".")[-1] != "progress") for k, v in sd.items()}
        main = [v for k, v in self.p.items() if not k.startswith(("se3_refine", "depth_scale")) and v.requires_grad]
        pose = [self.p["depth_scale.weight"], self.p["se3_refine.weight"]]
        self.opts = [torch.optim.Adam(main, lr=5e-4, eps=1e-8), torch.optim.Adam(pose, lr=2e-3, eps=1e-8)]
        self.sch = [torch.optim.lr_scheduler.ExponentialLR(self.opts[0], (5e-5 / 5e-4) ** (1 / max_steps)),
                    torch.optim.lr_scheduler.ExponentialLR(self.opts[1], (1e-5 / 2e-3) ** (1 / max_steps))]
        self.progress, self.step_no = 0.0, 0

    def sub(self, prefix):
        return {k[len(prefix) + 1:]: v for k, v in self.p.items() if k.startswith(prefix + ".")}

    def step(self, b, rng):
        p = self.p
        o, d = O.refine_and_cast(p["se3_refine.weight"], b["img_idx"], b["c2w"], b["directions"])
        rays = torch.cat([o, d, b["ray_infos"]], 1)
        depth = O.predicted_depth(p["depth_scale.weight"], b["img_idx"], b["inv_depths"], 0.1, 5.0)
        m = O.schedule_mult(self.progress)
        nets = {"nerf_coarse": self.sub("nerf_coarse"), "nerf_fine": self.sub("nerf_fine")}
        emb = {k: p[f"embedding_{k}.weight"] for k in ("coarse_a", "fine_a", "coarse_c", "fine_c")}
        res = O.render_rays(nets, self.cfgs, emb, rays, b["img_idx"], m, self.progress, N_samples=self.S, perturb=1.0,
                            N_importance=self.NI, rng=O.RenderRng(perturb_rand=rng["perturb_rand"], u=list(rng["u"])))
        if m > 0:
            O.blend_transient(res, O.transient_net(self.sub("transient_net"), b["feats"], b["img_idx"]))
        loss = sum(O.upnerf_loss(res, b["rgbs"], b["feats"], depth, m).values())
        for o_ in self.opts:
            o_.zero_grad()
        loss.backward()
        for o_, s_ in zip(self.opts, self.sch):
            o_.step()
            s_.step()
        self.step_no += 1
        # the reference stores progress in an fp32 Parameter and reads it back with .item()
        # (models/nerf_system.py:180,220-226): the schedule sees the fp32-ROUNDED value, e.g. 0.1 ->
        # 0.100000001 > candidate_schedule[0], which already selects phase 1 with sched_mult ~ 5e-17
        self.progress = float(torch.tensor((2 * self.step_no) / (2 * self.max_steps), dtype=torch.float32))
        return loss.detach(), res


def rng_for(R, S, NI, m, seed):
    ns = round(m * NI) if 0 < m < 1 else 0
    return dict(perturb_rand=synth.uniform((R, S), seed, 0, 1),
                u=[synth.uniform((R, NI - ns), seed + 1, 0, 1)] + ([synth.uniform((R, ns), seed + 2, 0, 1)] if ns else []))


class OracleOptimize:
    """The reference's test-time optimisation step (models/nerf_system_optmize.py:48-64, 84-150)
    restated on oracle functions: render with sched_mult = 1 / encode_candidate off, loss =
    mse(s_rgb_fine, rgbs); optimisers: Adam(5e-3) on the fresh embedding_fine_a + Adam(1e-4) on
    se3_refine with `pose_optimize`, else AdamW(1e-1) on embedding_fine_a (utils/optim.py:20-31)."""

    def __init__(self, cfgs, sd, S, NI, pose_optimize):
        # encode_candidate = False (:265-266) only drops the candidate inputs, which phase 2
        # (sched_mult = 1) never reads (models/nerf.py:96-104, models/rendering.py:152-219)
        self.cfgs = cfgs
        self.S, self.NI, self.pose_optimize = S, NI, pose_optimize
        self.p = {k: v.clone() for k, v in sd.items()}
        self.p["embedding_fine_a.weight"].requires_grad_(True)
        if pose_optimize:
            self.p["se3_refine.weight"].requires_grad_(True)
            self.opts = [torch.optim.Adam([self.p["embedding_fine_a.weight"]], lr=5e-3, eps=1e-8),
                         torch.optim.Adam([self.p["se3_refine.weight"]], lr=1e-4, eps=1e-8)]
        else:
            self.opts = [torch.optim.AdamW([self.p["embedding_fine_a.weight"]], lr=1e-1)]

    def sub(self, prefix):
        return {k[len(prefix) + 1:]: v for k, v in self.p.items() if k.startswith(prefix + ".")}

    def render(self, b, rng, perturb):
        p = self.p
        if self.pose_optimize:
            o, d = O.refine_and_cast(p["se3_refine.weight"], b["img_idx"], b["c2w"], b["directions"])
        else:
            o, d = O.get_rays(b["directions"], b["c2w"])
        rays = torch.cat([o, d, b["ray_infos"]], 1)
        nets = {"nerf_coarse": self.sub("nerf_coarse"), "nerf_fine": self.sub("nerf_fine")}
        emb = {k: p[f"embedding_{k}.weight"] for k in ("coarse_a", "fine_a", "coarse_c", "fine_c")}
        return O.render_rays(nets, self.cfgs, emb, rays, b["img_idx"], 1.0, float(p["nerf_coarse.progress"]),
                             N_samples=self.S, perturb=perturb, N_importance=self.NI,
                             rng=O.RenderRng(perturb_rand=rng["perturb_rand"], u=list(rng["u"])) if rng else None)

    def step(self, b, rng):
        res = self.render(b, rng, 1.0)
        loss = ((res["s_rgb_fine"] - b["rgbs"]) ** 2).mean()
        for o_ in self.opts:
            o_.zero_grad()
        loss.backward()
        for o_ in self.opts:
            o_.step()
        return loss.detach(), res
