"""metric / checkpoint mirrors against fixtures written from the real
utils/metric.py and utils/__init__.py (oracle/make_golden.py:gen_pose_metric).  The functions are
device-agnostic torch, so the CPU run pins the arithmetic and the GPU run repeats it on the device."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden"


def _load():
    z = np.load(GOLD / "pose_metric.npz")
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind != "U" else list(z[k])) for k in z.files}


def _check_pose_metric(dev):
    from upnerf_b200.utils import metric as M

    g = _load()
    est, gt = g["est"].to(dev), g["gt"].to(dev)
    parsed = M.parse_raw_camera(est)
    assert torch.allclose(parsed.cpu(), g["parsed"], atol=1e-6)
    err, aligned, gt_parsed = M.pose_metric(est, gt)
    assert torch.allclose(gt_parsed.cpu(), g["gt_parsed"], atol=1e-6)
    _, sim3 = M.prealign_cameras(parsed, gt_parsed)
    for k in ("R", "t0", "t1", "s0", "s1"):
        assert torch.allclose(sim3[k].cpu(), g[f"sim3_{k}"], atol=2e-6), k
    assert torch.allclose(aligned.cpu(), g["aligned"], atol=1e-5)
    # acos amplifies rounding near 0: 1e-4 rad on errors of ~0.07 rad
    assert torch.allclose(err["R"].cpu(), g["err_R"], atol=1e-4)
    assert torch.allclose(err["t"].cpu(), g["err_t"], atol=1e-5)
    assert torch.allclose(M.psnr(g["img_a"].to(dev), g["img_b"].to(dev)).cpu(), g["psnr"], atol=1e-5)
    assert torch.allclose(M.mse(g["img_a"].to(dev), g["img_b"].to(dev)).cpu(), g["mse"], atol=1e-7)


def test_pose_metric_cpu():
    _check_pose_metric(torch.device("cpu"))


def test_pose_metric_properties():
    """Procrustes alignment removes any global similarity: errors of exactly aligned sets vanish."""
    from upnerf_b200.utils import metric as M

    g = _load()
    gt = g["gt"]
    err, _, _ = M.pose_metric(gt.clone(), gt)
    assert float(err["R"].max()) < 1e-3 and float(err["t"].max()) < 1e-4
    # reflection branch: det(R) < 0 flips the last row
    X0 = torch.randn(16, 3, generator=torch.Generator().manual_seed(0))
    X1 = X0 * torch.tensor([1.0, 1.0, -1.0])
    s = M.procrustes_analysis(X0, X1)
    assert abs(float(torch.linalg.det(s.R)) - 1) < 1e-4 or abs(float(torch.linalg.det(s.R)) + 1) < 1e-4


def test_ssim_identity_and_range():
    from upnerf_b200.utils import metric as M

    a = torch.rand(1, 3, 16, 20, generator=torch.Generator().manual_seed(1))
    assert abs(float(M.ssim(a, a)) - 1) < 1e-6
    b = torch.rand(1, 3, 16, 20, generator=torch.Generator().manual_seed(2))
    v = float(M.ssim(a, b))
    assert -1 <= v < 0.5
    assert M.ssim(a, b, reduction="none").shape == a.shape


def test_extract_model_state_dict_matches_reference(tmp_path):
    from upnerf_b200.utils import extract_model_state_dict, load_ckpt

    g = _load()
    sd = {"nerf_coarse.xyz_encoding_1.0.weight": torch.ones(2), "nerf_coarse.progress": torch.zeros(1),
          "nerf_fine.xyz_encoding_1.0.weight": torch.ones(3), "embedding_fine_a.weight": torch.ones(1, 4),
          "se3_refine.weight": torch.zeros(2, 6), "transient_net.embedding_t.weight": torch.ones(1)}
    f = tmp_path / "x.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": {"max_steps": 10}}, f)
    for key in [k for k in g if k.startswith("ckptkeys__")]:
        _, name, n_ign = key.split("__")
        got = extract_model_state_dict(str(f), model_name=name, prefixes_to_ignore=["progress"] if int(n_ign) else [])
        assert sorted(got) == [str(x) for x in g[key]], key
    emb = torch.nn.Embedding(1, 4)
    load_ckpt(emb, str(f), model_name="embedding_fine_a")
    assert torch.equal(emb.weight.data, torch.ones(1, 4))
    # a bare state_dict file (no "state_dict" key) is accepted too (utils/__init__.py:7-8)
    torch.save(sd, f)
    assert sorted(extract_model_state_dict(str(f), "se3_refine")) == ["weight"]


@pytest.mark.gpu
def test_pose_metric_gpu(cuda_dev):
    _check_pose_metric(cuda_dev)


@pytest.mark.gpu
def test_checkpoint_round_trip_and_pose_eval(cuda_dev, tmp_path):
    """save_checkpoint writes what the reference's eval.py / tto.py read; load_checkpoint restores the
    parameters (views of the flat buffers stay valid), the step counter and the optimiser state, so a
    resumed run takes the same next step."""
    from oracle import synth
    from upnerf_b200.models.nerf_system import NeRFSystem
    from upnerf_b200.utils import ckpt as CK
    from upnerf_b200.utils import extract_model_state_dict

    hp = {"nerf.N_samples": 16, "nerf.N_importance": 16, "max_steps": 1000}
    n_img = 6

    def batch(seed):
        return {k: v.to(cuda_dev) for k, v in synth.ray_batch(64, n_img, seed).items()}

    torch.manual_seed(0)
    a = NeRFSystem(hp, N_images_train=n_img, device=cuda_dev)
    a.set_progress(0.3)
    a.global_step = 600
    for i in range(3):
        a.training_step(batch(i), i)
    f = tmp_path / "last.ckpt"
    ck = CK.save_checkpoint(a, str(f))
    assert {"state_dict", "hyper_parameters", "global_step"} <= set(ck)
    assert "se3_refine.weight" in ck["state_dict"] and "nerf_fine.xyz_encoding_final.weight" in ck["state_dict"]
    # the reference's per-module extraction works on the file
    sub = extract_model_state_dict(str(f), model_name="nerf_coarse")
    assert "xyz_encoding_1.0.weight" in sub and "progress" in sub

    torch.manual_seed(1)
    b = NeRFSystem(hp, N_images_train=n_img, device=cuda_dev)
    CK.load_checkpoint(b, str(f))
    assert b.global_step == a.global_step and abs(b._progress - a._progress) < 1e-9
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    rng = {"perturb_rand": synth.uniform((64, 16), 1, 0, 1).to(cuda_dev),
           "u": [synth.uniform((64, 8), 2, 0, 1).to(cuda_dev), synth.uniform((64, 8), 3, 0, 1).to(cuda_dev)]}
    la = a.training_step(batch(9), 3, rng=rng)
    lb = b.training_step(batch(9), 3, rng=rng)
    assert torch.allclose(la, lb, rtol=1e-5), (float(la), float(lb))
    wa, wb = a.group_main.flat.data, b.group_main.flat.data
    assert float((wa - wb).abs().max()) <= 1e-5 * float(wa.abs().max())

    # pose evaluation (eval.py:13-42): recover a known refinement
    g = _load()
    gt = g["gt"][:n_img].to(cuda_dev)
    from upnerf_b200.utils.camera import lie, pose
    true_refine = 0.05 * synth.uniform((n_img, 6), 5, -1, 1).to(cuda_dev)
    noised = pose.compose([pose_inv(lie.se3_to_SE3(true_refine)), gt])
    state = {"state_dict": {"se3_refine.weight": true_refine.cpu()}}
    r_deg, t_err, _ = CK.evaluate_poses(state, noised, gt)
    assert r_deg < 0.05 and t_err < 1e-3, (r_deg, t_err)
    zero = {"state_dict": {"se3_refine.weight": torch.zeros(n_img, 6)}}
    r0, t0, _ = CK.evaluate_poses(zero, noised, gt)
    assert r0 > 10 * max(r_deg, 1e-3)


def pose_inv(p):
    from upnerf_b200.utils.metric import _invert

    return _invert(p)


@pytest.mark.gpu
def test_resume_from_reference_layout_checkpoint(cuda_dev, tmp_path):
    """A checkpoint whose optimiser states are in the REFERENCE's layout (two torch.optim.Adam over ~150
    tensors in the order of models/nerf_system.py:340-409 / utils/optim.py:7-17) resumes with its Adam
    moments, per-tensor step counts and decayed learning rates -- not with a silently reset optimiser."""
    import warnings

    from upnerf_b200.models.nerf_system import NeRFSystem, adam_class
    from upnerf_b200.utils import ckpt as CK

    hp = {"nerf.N_samples": 16, "nerf.N_importance": 16, "max_steps": 1000}
    n_img, iters = 6, 40
    torch.manual_seed(0)
    sys_ = NeRFSystem(hp, N_images_train=n_img, device=cuda_dev)
    order = [CK._reference_param_order(sys_, w) for w in (0, 1)]
    # what the reference would have saved: per-tensor Adam states + ExponentialLR states after `iters` steps;
    # phase-2-only tensors (candidate head) stopped earlier, `progress` never got a gradient
    g = torch.Generator().manual_seed(4)
    name_of = {id(p): k for k, p in sys_.named_parameters()}
    ostates, sstates, lrs = [], [], []
    for w, lr0, lr_end in ((0, 5e-4, 5e-5), (1, 2e-3, 1e-5)):
        gamma = (lr_end / lr0) ** (1 / hp["max_steps"])
        state = {}
        for i, p in enumerate(order[w]):
            nm = name_of[id(p)]
            if adam_class(nm) == "never":
                continue
            step = 25 if adam_class(nm) == "cand" else iters
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": torch.randn(p.shape, generator=g) * 1e-3,
                        "exp_avg_sq": torch.rand(p.shape, generator=g) * 1e-6}
        lr = lr0 * gamma ** iters
        lrs.append(lr)
        ostates.append({"state": state, "param_groups": [{"lr": lr, "initial_lr": lr0, "betas": (0.9, 0.999), "eps": 1e-8,
                                                          "weight_decay": 0, "params": list(range(len(order[w])))}]})
        sstates.append({"gamma": gamma, "base_lrs": [lr0], "last_epoch": iters, "_step_count": iters + 1,
                        "_get_lr_called_within_step": False, "_last_lr": [lr]})
    f = tmp_path / "ref.ckpt"
    torch.save({"state_dict": {k: v.cpu() for k, v in sys_.state_dict().items()}, "global_step": 2 * iters,
                "optimizer_states": ostates, "lr_schedulers": sstates}, f)
    with warnings.catch_warnings():
        warnings.simplefilter("error")          # a full restore must not warn
        CK.load_checkpoint(sys_, str(f))
    for w, (opt, grp) in enumerate(((sys_.optimizer, sys_.group_main), (sys_.optimizer_pose, sys_.group_pose))):
        assert abs(opt.param_groups[0]["lr"] - lrs[w]) < 1e-12
        st = opt.state[grp.flat]
        for i, p in enumerate(order[w]):
            off, n = grp.offsets[id(p)]
            ref = ostates[w]["state"].get(i)
            if ref is None:
                assert float(st["exp_avg"][off:off + n].abs().max()) == 0
                continue
            assert torch.equal(st["exp_avg"][off:off + n].cpu(), ref["exp_avg"].reshape(-1))
            assert torch.equal(st["exp_avg_sq"][off:off + n].cpu(), ref["exp_avg_sq"].reshape(-1))
    assert sys_.optimizer.class_steps["always"] == iters and sys_.optimizer.class_steps["cand"] == 25
    assert sys_.optimizer.class_steps["never"] == 0 and sys_.optimizer_pose.class_steps["cand"] == 25
    # the next scheduler step continues the decay from the restored lr
    sys_.scheduler.step()
    gamma = (5e-5 / 5e-4) ** (1 / hp["max_steps"])
    assert abs(sys_.optimizer.param_groups[0]["lr"] - lrs[0] * gamma) < 1e-12
    # a checkpoint WITHOUT optimiser states: warned, lr placed on the schedule
    f2 = tmp_path / "bare.ckpt"
    torch.save({"state_dict": {k: v.cpu() for k, v in sys_.state_dict().items()}, "global_step": 2 * iters}, f2)
    torch.manual_seed(1)
    fresh = NeRFSystem(hp, N_images_train=n_img, device=cuda_dev)
    with pytest.warns(UserWarning, match="moments restart"):
        CK.load_checkpoint(fresh, str(f2))
    assert abs(fresh.optimizer.param_groups[0]["lr"] - lrs[0]) < 1e-10
    fresh.scheduler.step()
    assert abs(fresh.optimizer.param_groups[0]["lr"] - lrs[0] * gamma) < 1e-10
