"""CPU: pin oracle/train_step.py (OracleSystem, OracleOptimize) against fixtures written by the REAL
`NeRFSystem.training_step` (reference models/nerf_system.py:150-229, :41-73) and the real
`NeRFSystemOptimize.training_step` (models/nerf_system_optmize.py:48-64,84-150) -- see
oracle/make_golden.py:gen_train_step / gen_tto_step.  Seven steps with max_steps = 10 cross
phase 0 -> 1 -> 2, so the schedule, the optimiser wiring (two Adam + two ExponentialLR, scheduler
order, per-tensor Adam state that starts when a tensor first gets a gradient), the progress
update and the loss terms are all checked against the reference's own output."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.make_golden import TRAIN_CASE, _train_state
from oracle.train_step import OracleOptimize, OracleSystem

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(GOLD / f"{name}.npz")
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind != "U" else z[k]) for k in z.files}


def rng_of(g, it):
    u = [g[f"s{it}__u0"]] + ([g[f"s{it}__u1"]] if f"s{it}__u1" in g else [])
    return dict(perturb_rand=g[f"s{it}__perturb_rand"], u=u)


@pytest.mark.parametrize("name,want_phases", [("train_step_real", {0, 1, 2}), ("train_step_real_p03", {1}),
                                              ("train_step_real_p06", {2})])
def test_oracle_system_matches_real_training_step(name, want_phases):
    g = load(name)
    case = {k: int(g[f"case__{k}"]) for k in TRAIN_CASE}
    assert {k: v for k, v in case.items() if k != "n_steps"} == {k: v for k, v in TRAIN_CASE.items() if k != "n_steps"}
    R, S, NI, n_img = case["R"], case["S"], case["NI"], case["n_img"]
    cfgs, sd0 = _train_state(case)
    start = float(g["start"])
    for k in ("nerf_coarse.progress", "nerf_fine.progress"):
        sd0[k] = torch.tensor(start)
    orc = OracleSystem(cfgs, sd0, n_img, S, NI, case["max_steps"])
    orc.progress, orc.step_no = float(torch.tensor(start)), int(round(start * case["max_steps"]))
    phases = set()
    for it in range(case["n_steps"]):
        b = synth.ray_batch(R, n_img, 100 + it)
        assert abs(orc.progress - float(g[f"s{it}__progress"])) < 1e-7, it
        from oracle import upnerf_oracle as O

        m = O.schedule_mult(orc.progress)
        assert abs(m - float(g[f"s{it}__sched_mult"])) < 1e-7, (it, m)
        phases.add(0 if m == 0 else (2 if m == 1 else 1))
        loss, _ = orc.step(b, rng_of(g, it))
        ref = float(g[f"s{it}__loss"])
        assert abs(float(loss) - ref) <= 2e-6 * max(1.0, abs(ref)), (it, float(loss), ref)
        assert int(g[f"s{it}__global_step"]) == 2 * orc.step_no
        assert abs(orc.progress - float(g[f"s{it}__progress_after"])) < 1e-7
        # learning rates after the two scheduler steps (get_learning_rate, utils/optim.py:47-49)
        assert abs(orc.opts[0].param_groups[0]["lr"] - float(g[f"s{it}__log__lr"])) < 1e-9
        assert abs(orc.opts[1].param_groups[0]["lr"] - float(g[f"s{it}__log__lr_pose"])) < 1e-9
        # which tensors got no gradient in this phase, and the norms of the others
        gnone = set(g[f"s{it}__gnone"].tolist())
        for k, p in orc.p.items():
            if k.endswith("progress"):
                continue
            if k in gnone:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (it, k)
                continue
            ref_n = float(g[f"s{it}__gnorm__{k}"])
            got_n = float(p.grad.double().norm())
            if it == 0 and f"s{it}__gfull__{k}" in g:     # fresh state: element-wise, tight
                rg = g[f"s{it}__gfull__{k}"]
                assert float((p.grad - rg).norm()) <= 2e-4 * float(rg.norm()) + 1e-12, (it, k)
            # trajectories drift by Adam sign flips of noise-level gradients: 2e-4 on the first steps, 3e-3 later, 10 % from step 5 (per-ray pose gradients of 48 rays are chaotic; the fresh-start p03 / p06 fixtures pin those phases tightly)
            assert abs(got_n - ref_n) <= (2e-4 if it < 2 else (3e-3 if it < 5 else 0.1)) * max(ref_n, 1e-8) + 1e-9, (it, k, got_n, ref_n)
        # parameters after the step
        for k, p in orc.p.items():
            if k.endswith("progress"):
                continue
            v = p.detach()
            ref_n = float(g[f"s{it}__p__norm__{k}"])
            assert abs(float(v.double().norm()) - ref_n) <= (1e-5 if it < 1 else 1e-4) * max(ref_n, 1e-6), (it, k)
            if f"s{it}__p__full__{k}" in g:
                ref_v = g[f"s{it}__p__full__{k}"]
                upd_ref = ref_v - sd0[k]
                d = float((v - ref_v).norm())
                # Adam's early steps are ~lr*sign(g): elements whose gradient is rounding noise may
                # flip, so the bound is on the update norm-wise (measured: ~1e-4)
                assert d <= 2e-2 * float(upd_ref.norm()) + 1e-7, (it, k, d, float(upd_ref.norm()))
            head = g[f"s{it}__p__head__{k}"]
            got = v.reshape(-1)[: head.numel()]
            assert float((got - head).abs().max()) <= 2.5e-3 * (it + 1), (it, k)     # <= a few lr steps
    assert phases == want_phases


@pytest.mark.parametrize("tag", ["pose", "emb"])
def test_oracle_optimize_matches_real_tto_step(tag):
    g = load(f"tto_step_real_{tag}")
    case = {k: int(g[f"case__{k}"]) for k in TRAIN_CASE}
    R, S, NI, n_img = case["R"], case["S"], case["NI"], case["n_img"]
    cfgs, sd0 = _train_state(case)
    for k in ("nerf_coarse.progress", "nerf_fine.progress"):
        sd0[k] = torch.tensor(1.0)
    sd0["embedding_fine_a.weight"] = g["emb_fine_a0"].clone()
    orc = OracleOptimize(cfgs, sd0, S, NI, pose_optimize=(tag == "pose"))
    for it in range(case["n_steps"]):
        b = synth.ray_batch(R, n_img, 300 + it)
        loss, res = orc.step(b, rng_of(g, it))
        ref = float(g[f"s{it}__loss"])
        assert abs(float(loss) - ref) <= 2e-6 * max(1.0, abs(ref)), (it, float(loss), ref)
        ge = orc.p["embedding_fine_a.weight"].grad
        rg = g[f"s{it}__g_emb_fine_a"]
        assert float((ge - rg).norm()) <= 1e-4 * float(rg.norm()), it
        e = orc.p["embedding_fine_a.weight"].detach()
        upd = g[f"s{it}__emb_fine_a"] - g["emb_fine_a0"]
        assert float((e - g[f"s{it}__emb_fine_a"]).norm()) <= 1e-2 * float(upd.norm()), it
        se3 = orc.p["se3_refine.weight"].detach()
        if tag == "pose":
            rgp = g[f"s{it}__g_se3_refine"]
            assert float((orc.p["se3_refine.weight"].grad - rgp).norm()) <= 1e-4 * float(rgp.norm()), it
            upd = g[f"s{it}__se3_refine"] - sd0["se3_refine.weight"]
            assert float((se3 - g[f"s{it}__se3_refine"]).norm()) <= 1e-2 * float(upd.norm()), it
        else:
            assert torch.equal(se3, g[f"s{it}__se3_refine"])
