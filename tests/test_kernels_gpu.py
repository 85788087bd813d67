"""GPU parity of the individual hot-path kernels against the CPU oracle and the golden
fixtures produced by the real reference (tests/golden, oracle/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import synth
from oracle import upnerf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(GOLD / f"{name}.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def dev(x, d):
    return None if x is None else x.to(d).contiguous()


# ------------------------------------------------------------------ (a) pose + rays
def test_pose_rays_golden(cuda_dev):
    from upnerf_b200 import _lib as L

    g = load("pose_rays")
    R = g["img_idx"].shape[0]
    table, idx, c2w, dirs = (dev(g[k], cuda_dev) for k in ("table", "img_idx", "c2w", "directions"))
    nf = torch.tensor([[0.1, 5.0]]).repeat(R, 1).to(cuda_dev)
    rays = torch.empty(R, 8, device=cuda_dev)
    pose = torch.empty(R, 3, 4, device=cuda_dev)
    L.pose_rays_fwd(table, idx, c2w, dirs, nf, rays, pose)
    assert torch.allclose(pose.cpu(), g["refined"], atol=1e-6)
    assert torch.allclose(rays[:, :3].cpu(), g["rays_o"], atol=1e-6)
    assert torch.allclose(rays[:, 3:6].cpu(), g["rays_d"], atol=1e-6)
    assert torch.equal(rays[:, 6:].cpu(), nf.cpu())
    d_rays = torch.zeros(R, 8, device=cuda_dev)
    d_rays[:, :3], d_rays[:, 3:6] = g["co"].to(cuda_dev), g["cd"].to(cuda_dev)
    d_table = torch.zeros_like(table)
    L.pose_rays_bwd(table, idx, c2w, dirs, d_rays, d_table)
    ref = g["table_grad"]
    assert torch.isfinite(d_table).all()
    assert float((d_table.cpu() - ref).norm()) <= 1e-4 * float(ref.norm())
    # single-pose branch, no refinement (validation path)
    rays1 = torch.empty(R, 8, device=cuda_dev)
    L.pose_rays_fwd(None, None, c2w[0].contiguous(), dirs, nf, rays1)
    assert torch.allclose(rays1[:, :3].cpu(), g["single_o"], atol=1e-6)
    assert torch.allclose(rays1[:, 3:6].cpu(), g["single_d"], atol=1e-6)


def test_pose_rays_large_vs_oracle(cuda_dev):
    from upnerf_b200 import _lib as L

    R, n_img = 4096, 763
    b = synth.ray_batch(R, n_img, 17)
    table = synth.uniform((n_img, 6), 19, -0.05, 0.05)
    table[:50] = 0.0
    tab = table.clone().requires_grad_(True)
    o, d = O.refine_and_cast(tab, b["img_idx"], b["c2w"], b["directions"])
    co, cd = synth.uniform((R, 3), 23), synth.uniform((R, 3), 29)
    ((o * co).sum() + (d * cd).sum()).backward()
    rays = torch.empty(R, 8, device=cuda_dev)
    args = [dev(x, cuda_dev) for x in (table, b["img_idx"], b["c2w"], b["directions"])]
    L.pose_rays_fwd(*args, dev(b["ray_infos"], cuda_dev), rays)
    assert torch.allclose(rays[:, :3].cpu(), o.detach(), atol=1e-6)
    assert torch.allclose(rays[:, 3:6].cpu(), d.detach(), atol=1e-6)
    d_rays = torch.cat([co, cd, torch.zeros(R, 2)], 1).to(cuda_dev)
    d_table = torch.zeros(n_img, 6, device=cuda_dev)
    L.pose_rays_bwd(*args, d_rays, d_table)
    assert float((d_table.cpu() - tab.grad).norm()) <= 1e-4 * float(tab.grad.norm())


# ------------------------------------------------------------------ (b) positional encoding
@pytest.mark.parametrize("L_", [10, 4])
def test_posenc_golden(cuda_dev, L_):
    from upnerf_b200 import _lib as L

    g = load("posenc")
    x = dev(g["x"], cuda_dev)
    M = x.shape[0]
    for tag, c2f in (("none", None), ("p005", (0.1, 0.5)), ("p030", (0.1, 0.5)), ("p043", (0.1, 0.5)),
                     ("p075", (0.1, 0.5))):
        prog = g[f"prog_{tag}"].reshape(1).to(cuda_dev)
        bw = torch.empty(L_, device=cuda_dev)
        L.c2f_weights(prog, *(c2f or (0.0, 1.0)), c2f is not None, L_, bw)
        want_w = O.c2f_weights(L_, float(prog), c2f)
        assert torch.allclose(bw.cpu(), want_w, atol=1e-6)
        width = 3 + 6 * L_
        out = torch.full((M, width + 1), -7.0, device=cuda_dev)
        L.posenc_fwd(x, 3, M, L_, bw, out, width + 1, width + 1, L.F32)
        assert torch.allclose(out[:, :width].cpu(), g[f"L{L_}_{tag}"], atol=3e-6)
        assert (out[:, width] == 0).all()          # padding column is zeroed


def test_points_posenc_fwd_bwd_vs_oracle(cuda_dev):
    from upnerf_b200 import _lib as L

    R, S, Lx = 96, 64, 10
    rays = torch.cat([synth.uniform((R, 3), 1, -1, 1), torch.nn.functional.normalize(synth.uniform((R, 3), 2), dim=-1),
                      torch.full((R, 1), 0.1), torch.full((R, 1), 5.0)], 1)
    z = torch.sort(synth.uniform((R, S), 3, 0.1, 5.0), -1)[0]
    prog, c2f = 0.3, (0.1, 0.5)
    bw = O.c2f_weights(Lx, prog, c2f)
    rr = rays.clone().requires_grad_(True)
    xyz = rr[:, None, :3] + rr[:, None, 3:6] * z[..., None]
    pe = O.positional_encoding(xyz.reshape(-1, 3), Lx, prog, c2f)
    cot = synth.uniform(pe.shape, 4)
    (pe * cot).sum().backward()
    for dtype, tdt, tol in ((L.F32, torch.float32, 3e-6), (L.BF16, torch.bfloat16, 1e-2)):
        out = torch.empty(R * S, 64, device=cuda_dev, dtype=tdt)
        L.points_posenc_fwd(dev(rays, cuda_dev), dev(z, cuda_dev), Lx, dev(bw, cuda_dev), out, 64, 64, dtype)
        assert torch.allclose(out[:, :63].float().cpu(), pe.detach(), atol=tol, rtol=tol)
        assert (out[:, 63] == 0).all()
        d_pe = torch.zeros(R * S, 64, device=cuda_dev, dtype=tdt)
        d_pe[:, :63] = cot.to(cuda_dev).to(tdt)
        d_rays = torch.zeros(R, 8, device=cuda_dev)
        L.points_posenc_bwd(d_pe, 64, dev(rays, cuda_dev), dev(z, cuda_dev), Lx, dev(bw, cuda_dev), d_rays, dtype)
        ref = rr.grad[:, :6]
        rel = float((d_rays[:, :6].cpu() - ref).norm() / ref.norm())
        assert rel < (1e-5 if dtype == L.F32 else 1e-2), rel


# ------------------------------------------------------------------ (d) sampling
def test_stratified_z(cuda_dev):
    from upnerf_b200 import _lib as L

    R, S = 257, 64
    near = synth.uniform((R, 1), 5, 0.05, 0.5)
    far = synth.uniform((R, 1), 6, 3.0, 9.0)
    rays = torch.cat([torch.zeros(R, 6), near, far], 1)
    pr = synth.uniform((R, S), 7, 0, 1)
    for use_disp in (False, True):
        for perturb in (0.0, 1.0):
            want = O.stratified_z(near, far, S, use_disp, perturb, pr)
            z = torch.empty(R, S, device=cuda_dev)
            L.stratified_z(dev(rays, cuda_dev), dev(pr, cuda_dev), perturb, use_disp, S, z)
            # torch.linspace's rounding of the grid is implementation-defined (FMA or not): 1-ulp slack
            assert torch.allclose(z.cpu(), want, rtol=2e-6, atol=0)
            assert (z[:, 1:] >= z[:, :-1]).all()


def test_searchsorted_indices_bit_exact_golden(cuda_dev):
    """The index contract: identical CDFs and uniforms -> identical bin indices."""
    from upnerf_b200 import _lib as L

    g = load("sample_pdf")
    cdf, u = dev(g["cdf"], cuda_dev), dev(g["u"], cuda_dev)
    inds = torch.empty(u.shape, dtype=torch.int64, device=cuda_dev)
    L.searchsorted_right(cdf, u, inds)
    assert torch.equal(inds.cpu(), g["inds_u"])
    N = u.shape[1]
    det_u = torch.linspace(0, 1, N).expand(u.shape[0], N).contiguous().to(cuda_dev)
    L.searchsorted_right(cdf, det_u, inds)
    assert torch.equal(inds.cpu(), g["inds_det"])


def test_sample_pdf_golden(cuda_dev):
    from upnerf_b200 import _lib as L

    g = load("sample_pdf")
    bins, w = dev(g["bins"], cuda_dev), dev(g["weights"], cuda_dev)
    R, N = g["u"].shape
    for u_key, s_key, i_key in (("u", "samples_u", "inds_u"), ("u_rand", "samples_rand", None), (None, "samples_det", "inds_det")):
        u = dev(g[u_key], cuda_dev) if u_key else None
        out = torch.empty(R, N, device=cuda_dev)
        inds = torch.empty(R, N, dtype=torch.int64, device=cuda_dev)
        cdf = torch.empty(R, w.shape[1] + 1, device=cuda_dev)
        L.sample_pdf(bins, w, u, N, 1e-5, out, inds, cdf)
        # the kernel's own CDF may differ from torch's by summation order (1 ulp); given ITS cdf
        # the indices must be exactly torch.searchsorted's
        uu = g[u_key] if u_key else torch.linspace(0, 1, N).expand(R, N).contiguous()
        assert torch.equal(inds.cpu(), torch.searchsorted(cdf.cpu(), uu.contiguous(), right=True))
        assert torch.allclose(cdf.cpu(), g["cdf"], atol=2e-7)
        # ... and the samples must be the reference formula (rendering.py:33-49) evaluated on it
        c, ii, nw = cdf.cpu(), inds.cpu(), w.shape[1]
        below, above = (ii - 1).clamp_min(0), ii.clamp_max(nw)
        c0, c1 = c.gather(1, below), c.gather(1, above)
        b0, b1 = g["bins"].gather(1, below), g["bins"].gather(1, above)
        den = c1 - c0
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        assert torch.allclose(out.cpu(), b0 + (uu - c0) / den * (b1 - b0), rtol=1e-6, atol=1e-6)
        # against the reference's own output: (u - cdf)/denom amplifies the 1-ulp CDF difference in
        # near-empty bins (denom ~ 1e-5), so the bound scales with the conditioning of the bin:
        # |d sample| <= 3 |d cdf| (b1-b0)/denom with |d cdf| <= 2 ulp(1) = 2.4e-7, plus fp32 rounding
        # (a 1-ulp CDF difference can also move u across a bin edge: then the reference evaluated
        # the NEIGHBOURING bin, so take the worse conditioning of the two)
        err = (out.cpu() - g[s_key]).abs()
        ir = torch.searchsorted(g["cdf"], uu.contiguous(), right=True)
        rb, ra = (ir - 1).clamp_min(0), ir.clamp_max(nw)
        rden = g["cdf"].gather(1, ra) - g["cdf"].gather(1, rb)
        rden = torch.where(rden < 1e-5, torch.ones_like(rden), rden)
        cond = torch.maximum((b1 - b0).abs() / den, (g["bins"].gather(1, ra) - g["bins"].gather(1, rb)).abs() / rden)
        bound = 2e-5 + 4e-6 * cond
        assert (err <= bound).all(), float((err - bound).max())
        if i_key:
            # vs the reference's indices on ITS cdf: only ties can differ -- u = 1.0 of the det linspace
            # against cdf[-1] = 1 +- 1 ulp is one per ray (the sample is the last bin either way)
            assert (inds.cpu() != g[i_key]).float().mean() <= 1.0 / N + 1e-6


def test_sample_pdf_module_api(cuda_dev):
    from upnerf_b200.models.rendering import sample_pdf

    g = load("sample_pdf")
    out = sample_pdf(dev(g["bins"], cuda_dev), dev(g["weights"], cuda_dev), g["u"].shape[1], det=True)
    err = (out.cpu() - g["samples_det"]).abs()
    assert float(err.median()) < 1e-6 and float(err.max()) < 5e-3      # see test_sample_pdf_golden on conditioning


def test_resample_merge_large(cuda_dev):
    from upnerf_b200 import _lib as L

    R, S, N = 4096 + 3, 64, 64
    z = torch.sort(synth.uniform((R, S), 11, 0.1, 5.0), -1)[0]
    w_c = synth.uniform((R, S), 12, 0, 1) ** 3
    w_s = synth.uniform((R, S), 13, 0, 1) ** 3
    u0, u1 = synth.uniform((R, 40), 14, 0, 1), synth.uniform((R, 24), 15, 0, 1)
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    a = O.sample_pdf(mid, w_c[:, 1:-1], 40, u=u0)
    b = O.sample_pdf(mid, w_s[:, 1:-1], 24, u=u1)
    want = torch.sort(torch.cat([z, b, a], -1), -1)[0]
    out = torch.empty(R, S + N, device=cuda_dev)
    wc, ws = dev(w_c, cuda_dev), dev(w_s, cuda_dev)
    L.resample_merge(dev(z, cuda_dev), wc[:, 1:], ws[:, 1:], S, dev(u0, cuda_dev), dev(u1, cuda_dev), 40, 24, 1e-5, out)
    got = out.cpu()
    assert (got[:, 1:] >= got[:, :-1]).all()          # sortedness (size-independent property)
    err = (got - want).abs()
    assert float((err > 1e-4).float().mean()) < 1e-3  # ill-conditioned near-empty bins / a flipped bin, rarely
    assert float(err.median()) < 1e-6 and float(err.max()) < 0.2
    # deterministic single draw
    want = torch.sort(torch.cat([z, O.sample_pdf(mid, w_s[:, 1:-1], N, det=True)], -1), -1)[0]
    L.resample_merge(dev(z, cuda_dev), ws[:, 1:], None, S, None, None, N, 0, 1e-5, out)
    assert float(((out.cpu() - want).abs() > 1e-4).float().mean()) < 1e-3


# ------------------------------------------------------------------ (c) compositing
def _composite_case(R, S, seed):
    z = torch.sort(synth.uniform((R, S), seed, 0.1, 5.0), -1)[0]
    ss = synth.uniform((R, S), seed + 1, 0, 3) ** 2
    cs = synth.uniform((R, S), seed + 2, 0, 3) ** 2
    ss[0] = 0.0                       # empty ray
    ss[1, 3] = 500.0                  # saturated interior sample (alpha == 1)
    rgb = synth.uniform((R, S, 3), seed + 3, 0, 1)
    hf = synth.uniform((R, S, 256), seed + 4, -1, 1)
    g2 = synth.uniform((R, S, 128), seed + 5, -1, 1).clamp_min(0)
    return z, ss, cs, rgb, hf, g2


@pytest.mark.parametrize("mode", ["phase0", "phase1", "phase2", "nocand0"])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("S", [64, 61])      # 61: ragged last 32-sample chunk and load batch
def test_composite_fwd_bwd_vs_oracle(cuda_dev, mode, dtype, S):
    from upnerf_b200 import _lib as L

    R, F = 37, 24
    z, ss, cs, rgb, hf, g2 = _composite_case(R, S, 100)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    code = L.F32 if dtype == "f32" else L.BF16
    hf, g2 = hf.to(tdt).float(), g2.to(tdt).float()          # oracle sees the same rounded inputs
    m = {"phase0": 0, "phase1": 0.5, "phase2": 1, "nocand0": 0}[mode]
    cand = mode in ("phase0", "phase1")
    Wsf, bsf = synth.uniform((F, 256), 7, -0.1, 0.1), synth.uniform((F,), 8)
    Wcf, bcf = synth.uniform((F, 128), 9, -0.1, 0.1), synth.uniform((F,), 10)
    wcs = synth.uniform((128,), 11, -0.2, 0.2)
    leaves = [t.clone().requires_grad_(True) for t in (ss, cs, rgb, hf, g2)]
    ss_, cs_, rgb_, hf_, g2_ = leaves
    net_out = {"s_sigma": ss_, "c_sigma": cs_, "s_rgb": rgb_, "s_feat": hf_ @ Wsf.T + bsf, "c_feat": g2_ @ Wcf.T + bcf}
    res = {}
    O.composite(res, "x", net_out, z, m, cand, True)
    cots = {k: synth.uniform(v.shape, 200 + i) for i, (k, v) in enumerate(sorted(res.items()))}
    sum((res[k] * cots[k]).sum() for k in res).backward()

    a = L.CompositeArgs()
    a.R, a.S, a.cand, a.stat_rgb, a.dtype = R, S, int(cand), int(m > 0), code
    a.feat_mode = 0 if m == 1 else (2 if cand else 1)
    D = lambda t: t.to(cuda_dev).contiguous()
    t_in = dict(z=D(z), s_sigma=D(ss), c_sigma=D(cs), rgb=D(rgb.reshape(-1, 3)), hf=D(hf.reshape(-1, 256)).to(tdt),
                g2=D(g2.reshape(-1, 128)).to(tdt))
    for k, v in t_in.items():
        setattr(a, k, v.data_ptr())
    a.ld_hf, a.ld_g2 = 256, 128
    outs = dict(c_weights=(R, S), s_weights=(R, S), c_depth=(R,), t_weight=(R,), s_depth=(R,), s_rgb=(R, 3),
                hf_ray=(R, 256), g2_ray=(R, 128), ws_sum=(R,), wc_sum=(R,))
    t_out = {k: torch.full(s, float("nan"), device=cuda_dev) for k, s in outs.items()}
    for k, v in t_out.items():
        setattr(a, k, v.data_ptr())
    L.composite_fwd(a)
    torch.cuda.synchronize()
    tol = 2e-5
    got = {"s_depth_x": t_out["s_depth"]}
    if cand:
        got.update({"c_weights_x": t_out["c_weights"], "c_depth_x": t_out["c_depth"], "t_weight_x": t_out["t_weight"]})
    if m > 0 or not cand:
        got["s_weights_x"] = t_out["s_weights"]
    if m > 0:
        got["s_rgb_x"] = t_out["s_rgb"]
    if m < 1:
        feat = t_out["hf_ray"].cpu() @ Wsf.T + t_out["ws_sum"].cpu()[:, None] * bsf
        if cand:
            feat = feat + t_out["g2_ray"].cpu() @ Wcf.T + t_out["wc_sum"].cpu()[:, None] * bcf
        got["feat_x"] = feat
    assert set(got) == set(res)
    for k in res:
        assert torch.allclose(got[k].cpu(), res[k].detach(), atol=tol, rtol=1e-4), k

    # backward: upstream grads in the kernel's parametrisation
    gfeat = cots.get("feat_x")
    g_in = {}
    if gfeat is not None:
        g_in["g_hf_ray"] = gfeat @ Wsf
        g_in["g_ws_sum"] = gfeat @ bsf
        if cand:
            g_in["g_g2_ray"] = gfeat @ Wcf
            g_in["g_wc_sum"] = gfeat @ bcf
    for k in ("c_weights", "s_weights", "c_depth", "t_weight", "s_depth", "s_rgb"):
        if f"{k}_x" in cots:
            g_in["g_" + k] = cots[f"{k}_x"]
    g_dev = {k: D(v.float()) for k, v in g_in.items()}
    for k, v in g_dev.items():
        setattr(a, k, v.data_ptr())
    wcs_d = D(wcs)
    a.w_csigma = wcs_d.data_ptr()
    d_out = dict(d_ssig_pre=torch.zeros(R * S, device=cuda_dev), d_csig_pre=torch.zeros(R * S, device=cuda_dev),
                 d_rgb=torch.zeros(R * S, 3, device=cuda_dev), d_hf=torch.zeros(R * S, 256, device=cuda_dev, dtype=tdt),
                 d_g2pre=torch.zeros(R * S, 128, device=cuda_dev, dtype=tdt))
    for k, v in d_out.items():
        setattr(a, k, v.data_ptr())
    a.ld_dhf, a.ld_dg2 = 256, 128
    L.composite_bwd(a)
    torch.cuda.synchronize()
    gtol = 1e-4 if dtype == "f32" else 1e-2

    def rel(x, y):
        return float((x - y).norm() / (y.norm() + 1e-12))

    # d sigma_pre = d sigma * softplus'(pre) with softplus' = 1 - exp(-sigma)
    assert rel(d_out["d_ssig_pre"].cpu().reshape(R, S), ss_.grad * (1 - torch.exp(-ss))) < gtol
    if cand:
        dcp_ref = cs_.grad * (1 - torch.exp(-cs))
        assert rel(d_out["d_csig_pre"].cpu().reshape(R, S), dcp_ref) < gtol
    if m > 0:
        assert rel(d_out["d_rgb"].cpu().reshape(R, S, 3), rgb_.grad) < gtol
    if m < 1:
        assert rel(d_out["d_hf"].float().cpu().reshape(R, S, 256), hf_.grad) < gtol
        if cand:
            want = (g2_.grad + dcp_ref[..., None] * wcs) * (g2 > 0)
            assert rel(d_out["d_g2pre"].float().cpu().reshape(R, S, 128), want) < gtol


# ------------------------------------------------------------------ reference-granularity pose API
def test_camera_and_ray_modules_match_oracle_with_grads(cuda_dev):
    """lie.se3_to_SE3 -> pose.compose -> get_rays exactly as NeRFSystem.training_step chains them
    (models/nerf_system.py:158-163), with autograd through all three, vs the oracle / golden."""
    from upnerf_b200.utils import camera as cam
    from upnerf_b200.utils import ray as ray_utils

    g = load("pose_rays")
    idx = g["img_idx"].to(cuda_dev)
    table = g["table"].to(cuda_dev).requires_grad_(True)
    c2w = g["c2w"].to(cuda_dev)
    dirs = g["directions"].to(cuda_dev)
    refine = cam.lie.se3_to_SE3(table[idx])
    assert torch.allclose(refine.cpu(), g["se3"], atol=1e-6)
    refined = cam.pose.compose([refine, c2w])
    assert torch.allclose(refined.detach().cpu(), g["refined"], atol=1e-6)
    o, d = ray_utils.get_rays(dirs, refined)
    assert torch.allclose(o.detach().cpu(), g["rays_o"], atol=1e-6)
    assert torch.allclose(d.detach().cpu(), g["rays_d"], atol=1e-6)
    ((o * g["co"].to(cuda_dev)).sum() + (d * g["cd"].to(cuda_dev)).sum()).backward()
    ref = g["table_grad"]
    assert float((table.grad.cpu() - ref).norm()) <= 1e-4 * float(ref.norm())
    # single-pose branch with a gradient on the shared pose
    p1 = g["c2w"][0].to(cuda_dev).requires_grad_(True)
    o1, d1 = ray_utils.get_rays(dirs, p1)
    assert torch.allclose(d1.detach().cpu(), g["single_d"], atol=1e-6)
    pc = g["c2w"][0].clone().requires_grad_(True)
    oo, dd = O.get_rays(g["directions"], pc)
    ((oo * g["co"]).sum() + (dd * g["cd"]).sum()).backward()
    ((o1 * g["co"].to(cuda_dev)).sum() + (d1 * g["cd"].to(cuda_dev)).sum()).backward()
    assert torch.allclose(p1.grad.cpu(), pc.grad, atol=1e-4, rtol=1e-4)
    # fused form used by the train step
    nf = torch.tensor([[0.1, 5.0]]).repeat(len(idx), 1).to(cuda_dev)
    t2 = g["table"].to(cuda_dev).requires_grad_(True)
    rays = ray_utils.refine_rays(t2, idx, c2w, dirs, nf)
    ((rays[:, :3] * g["co"].to(cuda_dev)).sum() + (rays[:, 3:6] * g["cd"].to(cuda_dev)).sum()).backward()
    assert float((t2.grad.cpu() - ref).norm()) <= 1e-4 * float(ref.norm())


def test_get_ray_directions(cuda_dev):
    from upnerf_b200.utils.ray import get_ray_directions

    K = torch.tensor([[400.0, 0, 256], [0, 400.0, 192], [0, 0, 1]])
    d = get_ray_directions(6, 8, K)
    assert d.shape == (6, 8, 3)
    assert torch.allclose(d[2, 3], torch.tensor([(3 - 256) / 400.0, -(2 - 192) / 400.0, -1.0]))


# ------------------------------------------------------------------ (f3) fused Adam
def test_flat_adam_matches_per_tensor_torch_adam(cuda_dev):
    """upnerf_adam_step on one flat buffer with per-class liveness against torch.optim.Adam on the
    separate tensors, where a dead class has .grad None that step (skipped: no update, no step
    increment).  Segment boundaries are deliberately not multiples of 4."""
    from upnerf_b200.optim import FlatAdam

    sizes = [1, 1003, 64, 7, 2050, 5, 129]
    keys = ["never", "always", "rgb", "rgb", "cand", "always", "cand"]
    g = torch.Generator().manual_seed(5)
    init = [torch.randn(n, generator=g) for n in sizes]
    ref_p = [t.clone().to(cuda_dev).requires_grad_(True) for t in init]
    ref = torch.optim.Adam(ref_p, lr=3e-3, eps=1e-8, foreach=False, fused=False)
    flat = torch.nn.Parameter(torch.cat(init).to(cuda_dev))
    flat.grad = torch.zeros_like(flat)
    ends, off = [], 0
    for n, k in zip(sizes, keys):
        off += n
        ends.append((off, k))
    opt = FlatAdam(flat, ends, lr=3e-3, eps=1e-8)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.9)
    sched_ref = torch.optim.lr_scheduler.ExponentialLR(ref, 0.9)
    phases = [dict(rgb=False, cand=True), dict(rgb=False, cand=True), dict(rgb=True, cand=True),
              dict(rgb=True, cand=True), dict(rgb=True, cand=False), dict(rgb=True, cand=False)]
    for it, ph in enumerate(phases):
        live = {"always": True, "never": False, **ph}
        grads = [torch.randn(n, generator=g) * (10.0 ** (it % 3 - 1)) for n in sizes]
        flat.grad.zero_()
        off = 0
        for p_, gr, n, k in zip(ref_p, grads, sizes, keys):
            p_.grad = gr.to(cuda_dev) if live[k] else None
            if live[k]:
                flat.grad[off:off + n] = gr.to(cuda_dev)
            off += n
        ref.step()
        sched_ref.step()
        opt.set_live(live)
        opt.step()
        sched.step()
        want = torch.cat([p_.detach() for p_ in ref_p])
        err = float((flat.detach() - want).abs().max())
        assert err <= 2e-7, (it, err)
    assert opt.class_steps == {"never": 0, "always": 6, "rgb": 4, "cand": 4}
