"""Generate tests/golden/*.npz by running the REAL reference (mlvlab/UP-NeRF) on the CPU.

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):   python -m oracle.make_golden

The reference has no tests or golden vectors of its own, so these fixtures -- outputs of its
unmodified `models/rendering.py`, `models/nerf.py`, `utils/camera.py`, `utils/ray.py`,
`models/transient_net.py` and `losses.py` on small seeded inputs -- are what pins the oracle
(oracle/upnerf_oracle.py) and, through it or directly, the CUDA path.  Missing third-party
packages of the reference (easydict, kornia) are stubbed exactly as SURVEY.md Appendix A
describes; no reference source is copied or modified.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class EasyDict(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__ = self

    stub("easydict", EasyDict=EasyDict)

    def create_meshgrid(H, W, normalized_coordinates=False):
        xs, ys = torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H)
        return torch.stack(torch.meshgrid(xs, ys, indexing="ij"), -1).permute(1, 0, 2)[None]

    stub("kornia", create_meshgrid=create_meshgrid)
    import losses as ref_losses
    import models.nerf as ref_nerf
    import models.rendering as ref_rendering
    import models.transient_net as ref_tnet
    import utils.camera as ref_camera
    import utils.ray as ref_ray

    return ref_nerf, ref_rendering, ref_camera, ref_ray, ref_tnet, ref_losses


def _np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def _save(name, **arrays):
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **_np(arrays))
    size = (OUT / f"{name}.npz").stat().st_size
    print(f"  wrote {name}.npz ({size / 1024:.1f} KiB)")


def _ref_nerf_module(ref_nerf, cfg, sd):
    m = ref_nerf.NeRF(cfg.typ, D=cfg.D, W=cfg.W, skips=list(cfg.skips), encode_feat=cfg.encode_feat,
                      feat_dim=cfg.feat_dim, xyz_L=cfg.xyz_L, dir_L=cfg.dir_L,
                      appearance_dim=cfg.appearance_dim, candidate_dim=cfg.candidate_dim, c2f=cfg.c2f)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def _grad_digest(named_grads: dict) -> dict:
    """Store small gradients whole; for big ones the norm plus a leading block."""
    out = {}
    for k, g in named_grads.items():
        if g is None:
            out[f"gnone__{k}"] = np.zeros(0, np.float32)
            continue
        g = g.detach()
        out[f"gnorm__{k}"] = g.double().norm().float()
        if g.numel() <= 4096:
            out[f"gfull__{k}"] = g
        else:
            out[f"ghead__{k}"] = g[:4, :16].contiguous()
    return out


def gen_pose_rays(ref_camera, ref_ray):
    from . import synth

    n_img, R = 7, 12
    table = synth.uniform((n_img, 6), 11, -0.4, 0.4)
    table[0] = 0.0                         # theta = 0 row (poses start at identity)
    table[1, :3] *= 1e-4                   # tiny rotation
    table[2, :3] *= 6.0                    # large rotation (theta ~ 2)
    b = synth.ray_batch(R, n_img, 5)
    idx = b["img_idx"]
    idx[:3] = torch.tensor([0, 1, 2])
    table.requires_grad_(True)
    SE3 = ref_camera.lie.se3_to_SE3(table[idx])
    refined = ref_camera.pose.compose([SE3, b["c2w"]])
    o, d = ref_ray.get_rays(b["directions"], refined)
    co, cd = synth.uniform((R, 3), 21), synth.uniform((R, 3), 22)
    ((o * co).sum() + (d * cd).sum()).backward()
    # single-pose branch (validation path, utils/ray.py:57-65)
    o1, d1 = ref_ray.get_rays(b["directions"], b["c2w"][0])
    _save("pose_rays", table=table, img_idx=idx, c2w=b["c2w"], directions=b["directions"],
          se3=SE3, refined=refined, rays_o=o, rays_d=d, co=co, cd=cd, table_grad=table.grad,
          single_o=o1, single_d=d1)


def gen_posenc(ref_nerf):
    from . import synth
    from .upnerf_oracle import NerfConfig

    x = synth.uniform((10, 3), 31, -4, 4)
    arrays = {"x": x}
    for L in (10, 4):
        for tag, c2f, prog in (("none", None, 0.0), ("p005", (0.1, 0.5), 0.05),
                               ("p030", (0.1, 0.5), 0.30), ("p043", (0.1, 0.5), 0.43),
                               ("p075", (0.1, 0.5), 0.75)):
            cfg = NerfConfig(W=16, feat_dim=8, appearance_dim=0, candidate_dim=0, xyz_L=10, dir_L=4, c2f=c2f)
            m = ref_nerf.NeRF("coarse", W=16, feat_dim=8, xyz_L=10, dir_L=4, appearance_dim=0,
                              candidate_dim=0, c2f=c2f)
            m.progress.data.fill_(prog)
            arrays[f"L{L}_{tag}"] = m.positional_encoding(x, L)
            arrays[f"prog_{tag}"] = torch.tensor(prog)
    _save("posenc", **arrays)


def gen_sample_pdf(ref_rendering):
    from . import synth

    R, nb = 9, 15
    z = torch.sort(synth.uniform((R, nb + 1), 41, 0.1, 5.0), -1)[0]
    bins = 0.5 * (z[:, :-1] + z[:, 1:])
    w = synth.uniform((R, nb - 1), 42, 0, 1) ** 4
    w[0] = 0.0                       # all-zero weights -> uniform pdf from eps
    w[1, 3:] = 0.0                   # mass concentrated in the first bins
    w[2, :-1] = 0.0                  # mass in the last bin
    N = 12
    u = synth.uniform((R, N), 43, 0, 1)
    u[3, 0] = 0.0
    torch.manual_seed(1234)
    s_rand = ref_rendering.sample_pdf(bins, w, N, det=False)
    torch.manual_seed(1234)
    u_rand = torch.rand(R, N)
    s_det = ref_rendering.sample_pdf(bins, w, N, det=True)
    # explicit-u variant through the reference arithmetic (rendering.py:20-49) by patching rand
    orig = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        s_u = ref_rendering.sample_pdf(bins, w, N, det=False)
    finally:
        torch.rand = orig
    ww = w + 1e-5
    cdf = torch.cat([torch.zeros(R, 1), torch.cumsum(ww / ww.sum(-1, keepdim=True), -1)], -1)
    _save("sample_pdf", bins=bins, weights=w, u=u, u_rand=u_rand, samples_u=s_u, samples_rand=s_rand,
          samples_det=s_det, cdf=cdf, inds_u=torch.searchsorted(cdf, u.contiguous(), right=True),
          inds_det=torch.searchsorted(cdf, torch.linspace(0, 1, N).expand(R, N).contiguous(), right=True))


NET_CASES = {
    # name: (cfg kwargs, [(tag, sched_mult, progress)])
    "small": (dict(W=32, feat_dim=24, appearance_dim=6, candidate_dim=4, xyz_L=10, dir_L=4, c2f=(0.1, 0.5)),
              [("m0", 0, 0.05), ("m05", 0.5, 0.30), ("m1", 1, 0.75)]),
    "full": (dict(W=256, feat_dim=384, appearance_dim=48, candidate_dim=16, xyz_L=10, dir_L=4, c2f=(0.1, 0.5)),
             [("m0", 0, 0.05), ("m05", 0.5, 0.30), ("m03", 0.3, 0.22), ("m1", 1, 0.75)]),
    "c1": (dict(W=256, encode_feat=False, feat_dim=0, appearance_dim=0, candidate_dim=0, xyz_L=10, dir_L=4, c2f=None),
           [("m1", 1, 0.0)]),
}
NET_SEEDS = {"small": 3, "full": 5, "c1": 9}


def gen_nerf_forward(ref_nerf):
    from . import synth
    from .upnerf_oracle import NerfConfig

    M = 24
    for name, (kw, phases) in NET_CASES.items():
        cfg = NerfConfig(typ="coarse", **kw)
        xyz = synth.uniform((M, 3), 51, -2.5, 2.5)
        dirs = torch.nn.functional.normalize(synth.uniform((M, 3), 52), dim=-1)
        a = synth.uniform((M, max(cfg.appearance_dim, 1)), 53)[:, :cfg.appearance_dim]
        c = synth.uniform((M, max(cfg.candidate_dim, 1)), 54)[:, :cfg.candidate_dim]
        arrays = {"xyz": xyz, "dirs": dirs, "a": a, "c": c}
        for tag, m, prog in phases:
            sd = synth.nerf_state(cfg, NET_SEEDS[name], progress=prog)
            mod = _ref_nerf_module(ref_nerf, cfg, sd)
            inputs = {"input_xyz": xyz, "input_dir": dirs}
            if cfg.encode_appearance:
                inputs["input_a"] = a
            if cfg.encode_candidate:
                inputs["input_c"] = c
            out = mod(inputs, sched_mult=m)
            for k, v in out.items():
                arrays[f"{tag}__{k}"] = v
            arrays[f"{tag}__sched_mult"] = torch.tensor(float(m))
            arrays[f"{tag}__progress"] = torch.tensor(float(prog))
        _save(f"nerf_forward_{name}", **arrays)


def gen_render_rays(ref_nerf, ref_rendering):
    from . import synth
    from .upnerf_oracle import NerfConfig

    class Emb(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = torch.nn.Parameter(w.clone())

        def forward(self, idx):
            return self.weight[idx]

    n_img = 5
    for name, (kw, phases) in NET_CASES.items():
        R, S, NI = (6, 8, 8) if name != "c1" else (6, 8, 0)
        cfg_c, cfg_f = NerfConfig(typ="coarse", **kw), NerfConfig(typ="fine", **kw)
        b = synth.ray_batch(R, n_img, 61)
        from .upnerf_oracle import get_rays

        o, d = get_rays(b["directions"], b["c2w"])
        for tag, m, prog in phases:
            for mode, perturb in (("rand", 1.0), ("det", 0.0)):
                rays = torch.cat([o, d, b["ray_infos"]], 1).clone().requires_grad_(True)
                sd_c = synth.nerf_state(cfg_c, NET_SEEDS[name], progress=prog)
                sd_f = synth.nerf_state(cfg_f, NET_SEEDS[name] + 1, progress=prog)
                models = {"nerf_coarse": _ref_nerf_module(ref_nerf, cfg_c, sd_c)}
                if NI > 0:
                    models["nerf_fine"] = _ref_nerf_module(ref_nerf, cfg_f, sd_f)
                emb_w = synth.embeddings(n_img, cfg_c, 7)
                embs = {k: Emb(v) for k, v in emb_w.items()}
                seed = 777
                torch.manual_seed(seed)
                res = ref_rendering.render_rays(models=models, embeddings=embs, rays=rays, img_idx=b["img_idx"],
                                                sched_mult=m, sched_phase=0, N_samples=S, use_disp=False,
                                                perturb=perturb, N_importance=NI, white_back=False,
                                                encode_feat=cfg_c.encode_feat, validation=False)
                # replay the RNG stream in the reference's draw order
                arrays = {"rays": rays.detach(), "img_idx": b["img_idx"], "sched_mult": torch.tensor(float(m)),
                          "progress": torch.tensor(float(prog)), "perturb": torch.tensor(perturb),
                          "N_samples": S, "N_importance": NI, "n_img": n_img}
                if perturb > 0:
                    torch.manual_seed(seed)
                    arrays["perturb_rand"] = torch.rand(R, S)
                    if NI > 0:
                        if cfg_c.encode_candidate and 0 < m < 1:
                            ns = round(m * NI)
                            arrays["u0"] = torch.rand(R, NI - ns)
                            arrays["u1"] = torch.rand(R, ns)
                        else:
                            arrays["u0"] = torch.rand(R, NI)
                # scalar objective with fixed synthetic cotangents -> gradients
                loss = 0.0
                for j, (k, v) in enumerate(sorted(res.items())):
                    cot = synth.uniform(v.shape, 900 + j, -1, 1)
                    arrays[f"out__{k}"] = v
                    arrays[f"cot__{k}"] = cot
                    if "weights" in k:
                        continue        # used detached by sample_pdf; keep them out of the objective
                    loss = loss + (v * cot).sum()
                loss.backward()
                grads = {"rays": rays.grad}
                for ek, e in embs.items():
                    grads[f"emb_{ek}"] = e.weight.grad
                for mk, mod in models.items():
                    for pn, p in mod.named_parameters():
                        grads[f"{mk}.{pn}"] = p.grad
                arrays.update(_grad_digest(grads))
                _save(f"render_rays_{name}_{tag}_{mode}", **arrays)


def gen_tail(ref_tnet, ref_losses):
    from . import synth

    R, n_img = 16, 5
    torch.manual_seed(0)
    net = ref_tnet.TransientNet(N_images=n_img, beta_min=0.1, trasient_dim=128, feat_dim=384)
    sd = synth.transient_state(n_img, 3)
    assert list(sd) == list(net.state_dict())
    net.load_state_dict(sd)
    b = synth.ray_batch(R, n_img, 71)
    out = net(b["feats"], b["img_idx"])
    arrays = {"feats": b["feats"], "img_idx": b["img_idx"], "t_alpha": out["alpha"], "t_rgb": out["rgb"],
              "t_beta": out["beta"], "rgbs": b["rgbs"]}
    for m_tag, m in (("m0", 0), ("m05", 0.5), ("m1", 1)):
        res = {}
        for j, typ in enumerate(("coarse", "fine")):
            res[f"s_depth_{typ}"] = synth.uniform((R,), 80 + j, 0.2, 4.0)
            res[f"t_weight_{typ}"] = synth.uniform((R,), 82 + j, 0, 1)
            res[f"feat_{typ}"] = synth.uniform((R, 384), 84 + j, -0.1, 0.1)
            res[f"s_rgb_{typ}"] = synth.uniform((R, 3), 86 + j, 0, 1)
        res["t_beta"], res["t_alpha"] = out["beta"], out["alpha"]
        depth_t = synth.uniform((R,), 90, 0.2, 4.0)
        loss = ref_losses.UPNeRFLoss(depth_mult=1e-3, alpha_reg=1.0, encode_feat=True, fine=True)
        ld = loss(res, b["rgbs"], b["feats"], depth_t, m)
        for k, v in res.items():
            arrays[f"res__{k}"] = v
        arrays["depth_t"] = depth_t
        for k, v in ld.items():
            arrays[f"{m_tag}__{k}"] = v
    _save("tail", **arrays)


def gen_ray_batch():
    """`PhototourismDataset.__getitem__` (split "train", datasets/phototourism.py:420-454) + default_collate on
    synthetic tables.  The constructor reads a COLMAP scene from disk, so the object is made with
    `object.__new__` and given exactly the attributes `__getitem__` reads; the method itself is the
    reference's, unmodified."""
    from torch.utils.data import default_collate

    import datasets.phototourism as ref_ds

    from . import ray_batch as RB

    n_img, img_h, img_w, fh, C = 3, 7, 9, 5, 12
    t = RB.synth_tables(n_img, img_h, img_w, fh, C, seed=11)
    ds = object.__new__(ref_ds.PhototourismDataset)
    ds.split, ds.feat_map_dir = "train", "synthetic"
    ds.all_ray_infos, ds.all_directions, ds.all_rgbs = t["all_ray_infos"], t["all_directions"], t["all_rgbs"]
    ds.all_pxl_coords, ds.all_inv_depths, ds.feat_maps = t["all_pxl_coords"], t["all_inv_depths"], t["feat_maps"]
    ds.img_ids_train = [100 + 7 * i for i in range(n_img)]
    ds.poses_dict = {id_: t["poses"][i].numpy() for i, id_ in enumerate(ds.img_ids_train)}
    N = t["all_ray_infos"].shape[0]
    g = torch.Generator().manual_seed(5)
    # every ray of image 1 (all border cases: last row, last column, corner) + a shuffled sample
    idx = torch.cat([torch.arange(img_h * img_w, 2 * img_h * img_w), torch.randperm(N, generator=g)[:64]])
    batch = default_collate([ds[int(i)] for i in idx])
    arrays = {f"tab__{k}": v for k, v in t.items()}
    arrays["idx"] = idx
    for k, v in batch.items():
        arrays[f"out__{k}"] = v
    _save("ray_batch", **arrays)


def gen_pose_metric(ref_camera):
    """utils/metric.py:10-21,35-77 (psnr, pose_metric and its stages) and utils/__init__.py:4-19
    (extract_model_state_dict) from the real modules; `lpips` and `kornia.losses` (imported at module
    level by utils/metric.py, not used by these functions) are stubbed."""
    import tempfile

    lp = types.ModuleType("lpips")
    lp.LPIPS = lambda net="alex": None
    sys.modules["lpips"] = lp
    kl = types.ModuleType("kornia.losses")
    kl.ssim = types.SimpleNamespace(ssim_loss=None)
    sys.modules["kornia.losses"] = kl
    sys.modules["kornia"].losses = kl
    import utils as ref_utils
    import utils.metric as ref_metric

    g = torch.Generator().manual_seed(21)
    n = 24
    gt = ref_camera.lie.se3_to_SE3(torch.cat([0.6 * torch.randn(n, 3, generator=g), 2.0 * torch.randn(n, 3, generator=g)], 1))
    # estimated poses = GT moved by a global similarity-free rigid motion + per-camera noise
    glob = ref_camera.lie.se3_to_SE3(torch.tensor([[0.3, -0.2, 0.5, 1.0, -2.0, 0.5]]))[0]
    noise = ref_camera.lie.se3_to_SE3(0.03 * torch.randn(n, 6, generator=g))
    est = ref_camera.pose.compose([noise, gt, glob])
    err, aligned, gt_parsed = ref_metric.pose_metric(est, gt)
    assert err is not None
    parsed = torch.stack([ref_metric.parse_raw_camera(p) for p in est], 0)
    _, sim3 = ref_metric.prealign_cameras(parsed, gt_parsed)
    img_a, img_b = torch.rand(1, 3, 12, 16, generator=g), torch.rand(1, 3, 12, 16, generator=g)
    arrays = dict(est=est, gt=gt, err_R=err["R"], err_t=err["t"], aligned=aligned, gt_parsed=gt_parsed, parsed=parsed,
                  sim3_R=sim3.R, sim3_t0=sim3.t0, sim3_t1=sim3.t1, sim3_s0=sim3.s0, sim3_s1=sim3.s1,
                  img_a=img_a, img_b=img_b, psnr=ref_metric.psnr(img_a, img_b), mse=ref_metric.mse(img_a, img_b))
    # extract_model_state_dict on a Lightning-format file
    sd = {"nerf_coarse.xyz_encoding_1.0.weight": torch.ones(2), "nerf_coarse.progress": torch.zeros(1),
          "nerf_fine.xyz_encoding_1.0.weight": torch.ones(3), "embedding_fine_a.weight": torch.ones(4),
          "se3_refine.weight": torch.zeros(2, 6), "transient_net.embedding_t.weight": torch.ones(1)}
    with tempfile.TemporaryDirectory() as d:
        f = d + "/x.ckpt"
        torch.save({"state_dict": sd, "hyper_parameters": {"max_steps": 10}}, f)
        for name, ign in (("nerf_coarse", []), ("nerf_coarse", ["progress"]), ("embedding_fine_a", []), ("se3_refine", [])):
            got = ref_utils.extract_model_state_dict(f, model_name=name, prefixes_to_ignore=ign)
            arrays[f"ckptkeys__{name}__{len(ign)}"] = np.array(sorted(got), dtype="U64")
    _save("pose_metric", **arrays)


# ---------------------------------------------------------------------------------------------
# The REAL NeRFSystem.training_step / NeRFSystemOptimize.training_step
# ---------------------------------------------------------------------------------------------
TRAIN_CASE = dict(R=48, S=12, NI=12, n_img=6, max_steps=10, n_steps=7, seed=11)


def _import_systems():
    """models/nerf_system.py + models/nerf_system_optmize.py under the stubs of SURVEY.md Appendix A:
    absent third-party packages (pytorch_lightning, lpips, kornia.losses, matplotlib, imageio,
    mpl_toolkits) become empty modules; `LightningModule` is a minimal manual-optimisation shim whose
    `global_step` counts optimizer steps the way Lightning's manual-optimisation loop does (one per
    `optimizers()[i].step()`), which is what `progress = global_step / (2 * max_steps)`
    (models/nerf_system.py:220-226) relies on.  No reference source is modified."""

    def stub(name, **kw):
        if name in sys.modules and not kw:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    kl = stub("kornia.losses", ssim=types.SimpleNamespace(ssim_loss=None))
    sys.modules["kornia"].losses = kl
    stub("lpips", LPIPS=lambda net="alex": None)
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        stub(name)
    stub("mpl_toolkits.mplot3d.art3d", Poly3DCollection=None)

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.global_step = 0
            self.logged = {}

        def save_hyperparameters(self, h):
            self.hparams = dict(h)

        def log(self, k, v, **kw):
            self.logged[k] = v

        def manual_backward(self, loss):
            loss.backward()

        def optimizers(self):
            return self._opts if len(self._opts) > 1 else self._opts[0]

        def lr_schedulers(self):
            return self._scheds if len(self._scheds) > 1 else self._scheds[0]

    stub("pytorch_lightning", LightningModule=LightningModule)
    stub("pytorch_lightning.utilities")
    stub("pytorch_lightning.utilities.types", EPOCH_OUTPUT=None)
    import configs.config as ref_config
    import models.nerf_system as ref_sys
    import models.nerf_system_optmize as ref_opt

    return ref_config, ref_sys, ref_opt


class _LightningOptimizer:
    """What `self.optimizers()[i]` is under Lightning: forwards to the torch optimiser and advances the
    module's global_step once per step() (manual optimisation)."""

    def __init__(self, system, optimizer):
        self.system, self.optimizer = system, optimizer
        self.param_groups = optimizer.param_groups

    def zero_grad(self):
        self.optimizer.zero_grad()

    def step(self):
        self.optimizer.step()
        self.system.global_step += 1


def _train_state(case):
    """Initial full-width state (same recipe as tests/test_train_step_gpu.py:make_system)."""
    from . import synth
    from .train_step import KW
    from .upnerf_oracle import NerfConfig

    n_img, seed = case["n_img"], case["seed"]
    cfgs = {"nerf_coarse": NerfConfig(typ="coarse", **KW), "nerf_fine": NerfConfig(typ="fine", **KW)}
    sd = {}
    for k, cfg in cfgs.items():
        for pn, v in synth.nerf_state(cfg, seed + (k == "nerf_fine")).items():
            sd[f"{k}.{pn}"] = v
    for k, v in synth.embeddings(n_img, cfgs["nerf_coarse"], seed).items():
        sd[f"embedding_{k}.weight"] = v * 0.3
    for pn, v in synth.transient_state(n_img, seed).items():
        sd[f"transient_net.{pn}"] = v
    sd["se3_refine.weight"] = synth.uniform((n_img, 6), seed + 5, -0.02, 0.02)
    sd["depth_scale.weight"] = synth.uniform((n_img, 2), seed + 6, -0.05, 0.05)
    return cfgs, sd


def _build_system(cls, ref_config, case, extra=None):
    hp = ref_config.get_from_path(REF + "/configs/brandenburg_gate.yaml")
    hp.update({"nerf.N_samples": case["S"], "nerf.N_importance": case["NI"], "max_steps": case["max_steps"],
               "debug": True})
    hp.update(extra or {})
    system = cls(hp)
    system.train_dataset = types.SimpleNamespace(N_images_train=case["n_img"], N_images_test=case["n_img"],
                                                 white_back=False)
    # model_setup() moves the two pose tables `.to("cuda")` (models/nerf_system.py:399-402); this
    # container has no GPU, so device moves are no-ops while it runs
    orig_to = torch.nn.Module.to
    torch.nn.Module.to = lambda self, *a, **k: self
    try:
        ref_sys_cls = [c for c in cls.__mro__ if c.__name__ == "NeRFSystem"][0]
        ref_sys_cls.model_setup(system)
    finally:
        torch.nn.Module.to = orig_to
    system.log_pose = lambda: None
    return system


def _wire(system):
    out = system.configure_optimizers()
    opts, scheds = out if isinstance(out, tuple) else (out, [])
    system._opts = [_LightningOptimizer(system, o) for o in opts]
    system._scheds = [s["scheduler"] for s in scheds]


def _replay_rng(seed, R, S, NI, m, cand=True):
    """The reference's draw order inside render_rays: coarse perturbation, then one
    torch.rand per sample_pdf call."""
    torch.manual_seed(seed)
    out = {"perturb_rand": torch.rand(R, S)}
    if cand and 0 < m < 1:
        ns = round(m * NI)
        out["u0"], out["u1"] = torch.rand(R, NI - ns), torch.rand(R, ns)
    else:
        out["u0"] = torch.rand(R, NI)
    return out


def _state_digest(prefix, sd, arrays):
    for k, v in sd.items():
        v = v.detach().float()
        arrays[f"{prefix}__norm__{k}"] = v.double().norm().float()
        flat = v.reshape(-1)
        arrays[f"{prefix}__head__{k}"] = flat[:64].clone()
        if flat.numel() <= 4096:
            arrays[f"{prefix}__full__{k}"] = v.clone()


def gen_train_step(name="train_step_real", start=0.0, n_steps=None):
    """Real `NeRFSystem.training_step` calls (models/nerf_system.py:150-229) with max_steps = 10: from
    `start` = 0 seven steps see progress = 0, .1, ..., .6 and cross phase 0 -> 1 -> 2
    (candidate_schedule [0.1, 0.5]); two Adam optimisers + two ExponentialLR schedulers from the real
    configure_optimizers (:41-73).  The short `start` = 0.3 / 0.6 runs give phase-1 / phase-2 steps from a
    fresh optimiser state (no accumulated Adam drift), so they can be compared tightly."""
    from . import synth

    ref_config, ref_sys, ref_opt = _import_systems()
    case = dict(TRAIN_CASE)
    if n_steps is not None:
        case["n_steps"] = n_steps
    R, S, NI, n_img = case["R"], case["S"], case["NI"], case["n_img"]
    cfgs, sd0 = _train_state(case)
    for k in ("nerf_coarse.progress", "nerf_fine.progress"):
        sd0[k] = torch.tensor(float(start))
    system = _build_system(ref_sys.NeRFSystem, ref_config, case)
    missing = system.load_state_dict(sd0, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    _wire(system)
    system.global_step = int(round(start * 2 * case["max_steps"]))
    arrays = {f"case__{k}": torch.tensor(v) for k, v in case.items()}
    arrays["start"] = torch.tensor(float(start))
    arrays["state_keys"] = np.array(list(system.state_dict().keys()), dtype="U80")
    for it in range(case["n_steps"]):
        b = synth.ray_batch(R, n_img, 100 + it)
        progress = float(system.nerf_coarse.progress.detach())
        m = system.get_schedule_mult(progress)
        seed = 5000 + it
        for k, v in _replay_rng(seed, R, S, NI, m).items():
            arrays[f"s{it}__{k}"] = v
        torch.manual_seed(seed)
        loss = system.training_step({k: v.clone() for k, v in b.items()}, it)
        arrays[f"s{it}__progress"] = torch.tensor(progress)
        arrays[f"s{it}__sched_mult"] = torch.tensor(float(m))
        arrays[f"s{it}__loss"] = loss.detach()
        for k, v in system.logged.items():
            if k.startswith("train/") or k.startswith("lr"):
                arrays[f"s{it}__log__{k}"] = torch.as_tensor(v).detach().float().reshape(-1)
        arrays[f"s{it}__global_step"] = torch.tensor(system.global_step)
        arrays[f"s{it}__progress_after"] = system.nerf_coarse.progress.detach().clone()
        # gradients of the step as left in .grad (None = tensor not reached in this phase)
        gnames = []
        for k, p in system.named_parameters():
            if p.grad is None:
                gnames.append(k)
            else:
                arrays[f"s{it}__gnorm__{k}"] = p.grad.double().norm().float()
                if p.grad.numel() <= 4096 and start > 0:
                    arrays[f"s{it}__gfull__{k}"] = p.grad.detach().clone()
        arrays[f"s{it}__gnone"] = np.array(gnames, dtype="U80")
        _state_digest(f"s{it}__p", {k: v for k, v in system.state_dict().items() if "progress" not in k}, arrays)
    _save(name, **arrays)


def gen_tto_step():
    """Real `NeRFSystemOptimize.training_step` / `forward` / `configure_optimizers`
    (models/nerf_system_optmize.py:48-64,84-150), both `pose_optimize` settings, three steps each.  Its
    `model_setup` (:253-332) needs a checkpoint file, a COLMAP scene and GT poses: the part of it that
    matters to the step (:257-266 -- fresh embedding_fine_a over the test images as the only trained
    module, encode_candidate off) is applied here on top of the real `NeRFSystem.model_setup`."""
    import tempfile

    from . import synth

    ref_config, ref_sys, ref_opt = _import_systems()
    case = dict(TRAIN_CASE, n_steps=3)
    R, S, NI, n_img = case["R"], case["S"], case["NI"], case["n_img"]
    for pose_optimize in (True, False):
        cfgs, sd0 = _train_state(case)
        for k in ("nerf_coarse.progress", "nerf_fine.progress"):
            sd0[k] = torch.tensor(1.0)        # a trained checkpoint: PE fully open
        with tempfile.TemporaryDirectory() as d:
            system = _build_system(ref_opt.NeRFSystemOptimize, ref_config, case,
                                   extra={"pose_optimize": pose_optimize, "out_dir": d, "optimize_num": 0,
                                          "ckpt_path": d + "/none.ckpt"})
        system.load_state_dict(sd0, strict=True)
        with torch.no_grad():
            system.embedding_fine_a = torch.nn.Embedding(n_img, system.hparams["nerf.appearance_dim"])
            system.embedding_fine_a.weight.copy_(synth.uniform((n_img, 48), 77, -0.5, 0.5))
            system.embeddings["fine_a"] = system.embedding_fine_a
            system.models_to_train = [system.embedding_fine_a]
            system.nerf_coarse.encode_candidate = False
            system.nerf_fine.encode_candidate = False
        _wire(system)
        tag = "pose" if pose_optimize else "emb"
        arrays = {f"case__{k}": torch.tensor(v) for k, v in case.items()}
        arrays["emb_fine_a0"] = system.embedding_fine_a.weight.detach().clone()
        before = {k: v.detach().clone() for k, v in system.state_dict().items()}
        for it in range(case["n_steps"]):
            b = synth.ray_batch(R, n_img, 300 + it)
            seed = 7000 + it
            for k, v in _replay_rng(seed, R, S, NI, 1.0, cand=False).items():
                arrays[f"s{it}__{k}"] = v
            torch.manual_seed(seed)
            loss = system.training_step({k: v.clone() for k, v in b.items()}, it)
            arrays[f"s{it}__loss"] = loss.detach()
            arrays[f"s{it}__psnr"] = torch.as_tensor(system.logged["train/psnr"]).detach().float()
            arrays[f"s{it}__emb_fine_a"] = system.embedding_fine_a.weight.detach().clone()
            arrays[f"s{it}__se3_refine"] = system.se3_refine.weight.detach().clone()
            arrays[f"s{it}__g_emb_fine_a"] = system.embedding_fine_a.weight.grad.detach().clone()
            if pose_optimize:
                arrays[f"s{it}__g_se3_refine"] = system.se3_refine.weight.grad.detach().clone()
        after = system.state_dict()
        frozen = [k for k in before if k not in ("embedding_fine_a.weight", "se3_refine.weight")]
        assert all(torch.equal(before[k], after[k]) for k in frozen)     # networks stay frozen
        if not pose_optimize:
            assert torch.equal(before["se3_refine.weight"], after["se3_refine.weight"])
        _save(f"tto_step_real_{tag}", **arrays)


def main():
    ref_nerf, ref_rendering, ref_camera, ref_ray, ref_tnet, ref_losses = _import_reference()
    torch.set_num_threads(4)
    print("generating golden fixtures from", REF)
    gen_pose_rays(ref_camera, ref_ray)
    gen_posenc(ref_nerf)
    gen_sample_pdf(ref_rendering)
    gen_nerf_forward(ref_nerf)
    gen_render_rays(ref_nerf, ref_rendering)
    gen_tail(ref_tnet, ref_losses)
    gen_ray_batch()
    gen_pose_metric(ref_camera)
    gen_train_step()
    gen_train_step("train_step_real_p03", start=0.3, n_steps=2)
    gen_train_step("train_step_real_p06", start=0.6, n_steps=2)
    gen_tto_step()


if __name__ == "__main__":
    main()
