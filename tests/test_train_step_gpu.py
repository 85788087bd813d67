"""GPU: NeRFSystem.training_step (pose refinement + render + TransientNet + UPNeRFLoss + two Adam
steps) against the same step restated with the CPU oracle; PSNR parity after a fixed number of
steps on a synthetic scene (BASELINE.json: within 0.1 dB)."""
import math

import pytest
import torch

from oracle import synth
from oracle import upnerf_oracle as O
from oracle.train_step import KW, OracleSystem, rng_for

pytestmark = pytest.mark.gpu


def make_system(n_img, S, NI, precision, max_steps, device, seed=11):
    from upnerf_b200.models.nerf_system import NeRFSystem

    hp = {"nerf.N_samples": S, "nerf.N_importance": NI, "max_steps": max_steps, "kernel.precision": precision}
    torch.manual_seed(0)
    sys_ = NeRFSystem(hp, N_images_train=n_img, device=device)
    cfgs = {"nerf_coarse": O.NerfConfig(typ="coarse", **KW), "nerf_fine": O.NerfConfig(typ="fine", **KW)}
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
    sys_.load_state_dict(sd)
    return sys_, cfgs, sd


@pytest.mark.parametrize("start_progress", [0.05, 0.3, 0.75])
def test_training_step_matches_oracle_fp32(cuda_dev, start_progress):
    R, S, NI, n_img, max_steps = 128, 32, 32, 12, 1000
    sys_, cfgs, sd = make_system(n_img, S, NI, "fp32", max_steps, cuda_dev)
    orc = OracleSystem(cfgs, sd, n_img, S, NI, max_steps)
    sys_.set_progress(start_progress)
    sys_.global_step = int(round(start_progress * 2 * max_steps))
    orc.progress, orc.step_no = start_progress, int(round(start_progress * max_steps))
    for it in range(2):
        b = synth.ray_batch(R, n_img, 100 + it)
        m = O.schedule_mult(orc.progress)
        rng = rng_for(R, S, NI, m, 200 + 10 * it)
        l_ref, _ = orc.step(b, rng)
        bd = {k: v.to(cuda_dev) for k, v in b.items()}
        l = sys_.training_step(bd, it, rng=rng)
        assert abs(float(l) - float(l_ref)) <= 1e-4 * max(1.0, abs(float(l_ref))), (it, float(l), float(l_ref))
    # parameters after two Adam steps
    # Adam's first steps are ~lr*sign(g): an element whose gradient is at rounding-noise level may
    # legitimately take the opposite +-lr step, so updates are compared norm-wise per tensor.
    own = sys_.state_dict()
    worst = 0.0
    for k, v in orc.p.items():
        if k.endswith("progress"):
            continue
        upd_ref = v.detach() - sd[k]
        upd = own[k].cpu() - sd[k]
        if float(upd_ref.abs().max()) == 0:
            assert float(upd.abs().max()) == 0, k
            continue
        r = float((upd - upd_ref).norm() / upd_ref.norm())
        worst = max(worst, r)
        assert r <= 0.1, (k, r)
    assert abs(sys_._progress - orc.progress) < 1e-9
    print(f"worst relative update error after 2 Adam steps: {worst:.3e}")


def test_psnr_parity_after_fixed_steps(cuda_dev):
    """Same synthetic scene, same seeds: PSNR of the fp32 and bf16 CUDA paths vs the oracle after
    N steps in phase 2 (rgb only) must agree within 0.1 dB (BASELINE.json)."""
    R, S, NI, n_img, max_steps, n_steps = 256, 32, 32, 8, 40, 12
    psnrs = {}
    for which in ("oracle", "fp32", "bf16"):
        sys_, cfgs, sd = make_system(n_img, S, NI, "bf16" if which == "bf16" else "fp32", max_steps, cuda_dev)
        orc = OracleSystem(cfgs, sd, n_img, S, NI, max_steps) if which == "oracle" else None
        start = 0.75
        if orc:
            orc.progress, orc.step_no = start, int(round(start * max_steps))
        else:
            sys_.set_progress(start)
            sys_.global_step = int(round(start * 2 * max_steps))
        vals = []
        for it in range(n_steps):
            b = synth.ray_batch(R, n_img, 300 + it)
            # a learnable target: colour is a smooth function of the ray direction
            b["rgbs"] = 0.5 + 0.5 * torch.sin(3.0 * b["directions"] + torch.tensor([0.0, 1.0, 2.0]))
            rng = rng_for(R, S, NI, 1, 400 + 10 * it)
            if orc:
                _, res = orc.step(b, rng)
                vals.append(float(O.psnr(res["s_rgb_fine"].detach(), b["rgbs"])))
            else:
                sys_.training_step({k: v.to(cuda_dev) for k, v in b.items()}, it, rng=rng)
                vals.append(float(sys_.logged["train/psnr"]))
        psnrs[which] = vals
    print({k: [round(x, 3) for x in v[-3:]] for k, v in psnrs.items()})
    assert psnrs["oracle"][-1] > psnrs["oracle"][0] + 0.5            # it actually trains
    for which in ("fp32", "bf16"):
        assert abs(psnrs[which][-1] - psnrs["oracle"][-1]) <= 0.1, (which, psnrs[which][-1], psnrs["oracle"][-1])
        assert math.isfinite(psnrs[which][-1])


@pytest.mark.parametrize("m", [0.0, 0.37, 1.0])
@pytest.mark.parametrize("cand", [True, False])
def test_fused_tail_matches_torch_loss(cuda_dev, m, cand):
    """upnerf_tail_loss (depth correction + UPNeRFLoss + backward + psnr in one launch) against the
    torch module `UPNeRFLoss` (reference losses.py:21-64, nerf_system.py:169-177) + autograd."""
    from upnerf_b200 import _lib as L
    from upnerf_b200.losses import UPNeRFLoss, fused_tail

    R, F, n_img = 1000, 384, 9
    g = torch.Generator().manual_seed(int(m * 100) + cand)
    rnd = lambda *s: torch.rand(*s, generator=g)
    batch = {"rgbs": rnd(R, 3), "feats": torch.nn.functional.normalize(torch.randn(R, F, generator=g), dim=-1),
             "img_idx": torch.randint(0, n_img, (R,), generator=g), "inv_depths": 0.05 + 12 * rnd(R)}
    batch = {k: v.to(cuda_dev) for k, v in batch.items()}
    ds = (0.3 * torch.randn(n_img, 2, generator=g)).to(cuda_dev).requires_grad_(True)
    res = {}
    for typ in ("coarse", "fine"):
        res[f"s_depth_{typ}"] = (0.1 + 5 * rnd(R)).to(cuda_dev).requires_grad_(True)
        if m < 1:
            res[f"feat_{typ}"] = (0.1 * torch.randn(R, F, generator=g)).to(cuda_dev).requires_grad_(True)
            if cand:
                res[f"t_weight_{typ}"] = rnd(R).to(cuda_dev).requires_grad_(True)
        if m > 0:
            res[f"s_rgb_{typ}"] = rnd(R, 3).to(cuda_dev).requires_grad_(True)
    if m > 0:
        res["t_beta"] = (0.1 + rnd(R, 1)).to(cuda_dev).requires_grad_(True)
        res["t_alpha"] = rnd(R, 1).to(cuda_dev).requires_grad_(True)
    # torch reference
    scale, shift = torch.unbind(ds[batch["img_idx"]], 1)
    inv = batch["inv_depths"] * torch.exp(scale) + shift
    inv = torch.where(inv < 1 / 5.0, torch.full_like(inv, 1 / 5.0), inv)
    depth = 1.0 / inv
    depth = torch.where(depth < 0.1, torch.full_like(depth, 0.1), depth)
    assert 0 < int((depth == 0.1).sum()) and 0 < int((inv == 0.2).sum())       # both clamps are exercised
    ref_d = UPNeRFLoss(depth_mult=1e-3, alpha_reg=1.0)(res, batch["rgbs"], batch["feats"], depth, m)
    ref = sum(ref_d.values())
    ref.backward()
    ref_g = {k: v.grad.clone() for k, v in res.items() if v.grad is not None}
    ref_ds = ds.grad.clone() if ds.grad is not None else torch.zeros_like(ds)
    ds.grad = torch.zeros_like(ds)
    ws = torch.zeros(L.tail_workspace_bytes(), device=cuda_dev, dtype=torch.uint8)
    for rep in range(2):        # second call: the ticket re-arms itself
        ds.grad.zero_()
        losses, roots, grads = fused_tail(res, batch, ds, m, depth_mult=1e-3, alpha_reg=1.0, near=0.1, far=5.0,
                                          fine=True, workspace=ws)
        torch.cuda.synchronize()
        got = dict(zip(L.TAIL_TERMS, losses[:8].tolist()))
        for k, v in ref_d.items():
            assert abs(got[k] - float(v)) <= 1e-5 * max(1.0, abs(float(v))), (k, got[k], float(v))
        assert all(got[k] == 0.0 for k in got if k not in ref_d)
        assert abs(float(losses[8]) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        if m > 0:
            psnr = -10 * torch.log10(((res["s_rgb_fine"] - batch["rgbs"]) ** 2).mean())
            assert abs(float(losses[9]) - float(psnr)) <= 1e-4
        by_id = {id(r): gr for r, gr in zip(roots, grads)}
        for k, v in res.items():
            if k in ref_g:
                assert id(v) in by_id, k
                err = float((by_id[id(v)] - ref_g[k]).abs().max())
                assert err <= 1e-6 * max(1.0, float(ref_g[k].abs().max()) * 1e3), (k, err)
                assert torch.allclose(by_id[id(v)], ref_g[k], rtol=1e-4, atol=1e-9), k
            else:
                assert id(v) not in by_id, k      # t_weight is detached in the reference
        assert torch.allclose(ds.grad, ref_ds, rtol=1e-4, atol=1e-9)


def test_training_step_unfused_tail_path(cuda_dev):
    """`kernel.fused_tail=False` keeps the torch UPNeRFLoss path: same loss and update as the fused one."""
    R, S, NI, n_img, max_steps = 128, 32, 32, 12, 1000
    outs = []
    for fused in (True, False):
        sys_, cfgs, sd = make_system(n_img, S, NI, "fp32", max_steps, cuda_dev)
        sys_.hparams["kernel.fused_tail"] = fused
        sys_.set_progress(0.3)
        sys_.global_step = 600
        b = {k: v.to(cuda_dev) for k, v in synth.ray_batch(R, n_img, 100).items()}
        rng = rng_for(R, S, NI, O.schedule_mult(0.3), 200)
        l = sys_.training_step(b, 0, rng=rng)
        outs.append((float(l), sys_.group_main.flat.grad.clone(), sys_.group_pose.flat.grad.clone(),
                     float(sys_.logged["train/psnr"]), {k: float(v) for k, v in sys_.logged.items() if k.startswith("train/l_")}))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-5 * max(1.0, abs(outs[1][0]))
    assert abs(outs[0][3] - outs[1][3]) <= 1e-3
    assert outs[0][4].keys() == outs[1][4].keys()
    for i in (1, 2):
        assert float((outs[0][i] - outs[1][i]).norm() / outs[1][i].norm()) <= 1e-4


def test_training_steps_across_phase_boundary_fp32(cuda_dev):
    """Five steps that start in phase 0 and cross into phase 1: the rgb head, the appearance
    embeddings and TransientNet get their first gradient at step 3, so the reference's per-tensor
    Adam starts THEIR bias correction at t = 1 there (tensors with .grad None are skipped,
    models/nerf_system.py:188-195).  FlatAdam keeps one step counter per liveness class; a single
    global counter would be off by ~40 % on the first live update."""
    R, S, NI, n_img, max_steps = 128, 32, 32, 12, 20
    sys_, cfgs, sd = make_system(n_img, S, NI, "fp32", max_steps, cuda_dev)
    orc = OracleSystem(cfgs, sd, n_img, S, NI, max_steps)
    sys_.set_progress(0.0)
    sys_.global_step = 0
    ms = []
    for it in range(5):
        b = synth.ray_batch(R, n_img, 500 + it)
        m = O.schedule_mult(orc.progress)
        ms.append(m)
        rng = rng_for(R, S, NI, m, 600 + 10 * it)
        l_ref, _ = orc.step(b, rng)
        l = sys_.training_step({k: v.to(cuda_dev) for k, v in b.items()}, it, rng=rng)
        assert abs(float(l) - float(l_ref)) <= 2e-4 * max(1.0, abs(float(l_ref))), (it, float(l), float(l_ref))
    assert ms[0] == 0 and ms[-1] > 0, ms          # the run really crosses the boundary
    own = sys_.state_dict()
    worst = {}
    for k, v in orc.p.items():
        if k.endswith("progress"):
            continue
        upd_ref = v.detach() - sd[k]
        upd = own[k].cpu() - sd[k]
        if float(upd_ref.abs().max()) == 0:
            assert float(upd.abs().max()) == 0, k
            continue
        r = float((upd - upd_ref).norm() / upd_ref.norm())
        worst[k] = r
        assert r <= 0.1, (k, r)
    late = [k for k in worst if "rgb_share_layer" in k or k.startswith(("embedding_coarse_a", "transient_net.feat"))]
    assert late, "expected late-starting tensors in the comparison"
    print("worst late-start tensor:", max(worst[k] for k in late), "worst overall:", max(worst.values()))


@pytest.mark.parametrize("name", ["train_step_real", "train_step_real_p03", "train_step_real_p06"])
def test_training_step_matches_real_reference_golden(cuda_dev, name):
    """The CUDA train step (fp32 mode) against fixtures written by the REAL `NeRFSystem.training_step`
    (reference models/nerf_system.py:150-229; oracle/make_golden.py:gen_train_step) -- no oracle in
    between: loss and every logged loss term per step, the schedule (fp32-rounded progress, phases
    0 -> 1 -> 2), learning rates, which tensors get no gradient, gradient norms and parameters."""
    from pathlib import Path

    import numpy as np

    from oracle.make_golden import TRAIN_CASE, _train_state
    from upnerf_b200.models.nerf_system import NeRFSystem

    z = np.load(Path(__file__).resolve().parent / "golden" / f"{name}.npz")
    g = {k: (torch.from_numpy(z[k]) if z[k].dtype.kind != "U" else z[k]) for k in z.files}
    case = {k: int(g[f"case__{k}"]) for k in TRAIN_CASE}
    R, S, NI, n_img = case["R"], case["S"], case["NI"], case["n_img"]
    cfgs, sd0 = _train_state(case)
    start = float(g["start"])
    for k in ("nerf_coarse.progress", "nerf_fine.progress"):
        sd0[k] = torch.tensor(start)
    hp = {"nerf.N_samples": S, "nerf.N_importance": NI, "max_steps": case["max_steps"], "kernel.precision": "fp32"}
    sys_ = NeRFSystem(hp, N_images_train=n_img, device=cuda_dev)
    sys_.load_state_dict(sd0)
    sys_.global_step = int(round(start * 2 * case["max_steps"]))
    names = [k for k in sys_.state_dict() if not k.endswith("progress")]
    for it in range(case["n_steps"]):
        b = synth.ray_batch(R, n_img, 100 + it)
        assert abs(sys_._progress - float(g[f"s{it}__progress"])) < 1e-7, it
        m = sys_.get_schedule_mult(sys_._progress)
        assert abs(m - float(g[f"s{it}__sched_mult"])) < 1e-7, (it, m)
        rng = dict(perturb_rand=g[f"s{it}__perturb_rand"],
                   u=[g[f"s{it}__u0"]] + ([g[f"s{it}__u1"]] if f"s{it}__u1" in g else []))
        l = sys_.training_step({k: v.to(cuda_dev) for k, v in b.items()}, it, rng=rng)
        ref = float(g[f"s{it}__loss"])
        assert abs(float(l) - ref) <= 1e-4 * max(1.0, abs(ref)), (it, float(l), ref)
        for k, v in sys_.logged.items():
            if k.startswith("train/l_"):
                r = float(g[f"s{it}__log__{k}"])
                assert abs(float(v) - r) <= 1e-4 * max(1.0, abs(r)), (it, k, float(v), r)
        want_terms = {k[len(f"s{it}__log__"):] for k in g if k.startswith(f"s{it}__log__train/l_")}
        assert {k for k in sys_.logged if k.startswith("train/l_")} == want_terms, it
        assert abs(float(sys_.logged["train/psnr"]) - float(g[f"s{it}__log__train/psnr"])) <= 2e-3, it
        assert sys_.global_step == int(g[f"s{it}__global_step"])
        assert abs(sys_._progress - float(g[f"s{it}__progress_after"])) < 1e-7
        assert abs(sys_.optimizer.param_groups[0]["lr"] - float(g[f"s{it}__log__lr"])) < 1e-9
        assert abs(sys_.optimizer_pose.param_groups[0]["lr"] - float(g[f"s{it}__log__lr_pose"])) < 1e-9
        gnone = set(g[f"s{it}__gnone"].tolist())
        own = sys_.state_dict()
        params = dict(sys_.named_parameters())
        gtol = 1e-2 if it < 1 else (3e-2 if it < 5 else 0.15)       # Adam-trajectory drift, see test_train_step_golden.py
        for k in names:
            gr = params[k].grad
            if k in gnone:
                assert gr is None or float(gr.abs().max()) == 0.0, (it, k)
            else:
                ref_n = float(g[f"s{it}__gnorm__{k}"])
                got_n = float(gr.double().norm())
                assert abs(got_n - ref_n) <= gtol * max(ref_n, 1e-8) + 1e-9, (it, k, got_n, ref_n)
                if it == 0 and f"s{it}__gfull__{k}" in g:
                    rg = g[f"s{it}__gfull__{k}"]
                    assert float((gr.cpu() - rg).norm()) <= 1e-2 * float(rg.norm()) + 1e-12, (it, k)
            # parameters: Adam's first steps are ~lr * sign(g), so an element whose gradient is at rounding-noise
            # level may take the opposite +-lr step; everything is bounded relative to the size of the UPDATE
            v = own[k].cpu()
            lr_now = (sys_.optimizer_pose if k.startswith(("se3_refine", "depth_scale")) else sys_.optimizer).param_groups[0]["lr"]
            upd_bound = lr_now * (it + 1) * v.numel() ** 0.5 * 1.05          # every element moved by <= lr per step
            ref_n = float(g[f"s{it}__p__norm__{k}"])
            assert abs(float(v.double().norm()) - ref_n) <= 0.1 * upd_bound + 1e-7, (it, k)
            if f"s{it}__p__full__{k}" in g and it < 2:
                ref_v = g[f"s{it}__p__full__{k}"]
                upd_ref = ref_v - sd0[k]
                d = float((v - ref_v).norm())
                assert d <= 0.15 * float(upd_ref.norm()) + 1e-7, (it, k, d, float(upd_ref.norm()))
