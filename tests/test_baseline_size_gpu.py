"""GPU parity at BASELINE.json's benchmark size (config 2/3: 4096 rays x (64 + 64) samples, 763 images,
all heads + embeddings, phase 1) in both precisions: outputs and norm-wise parameter / embedding /
pose-path gradients of `render_rays` against the CPU oracle run in fp64 ("truth") and in fp32 (what
the reference's own fp32 arithmetic gives).  The oracle processes the batch in ray chunks; the
objective is linear in the outputs (fixed cotangents), so chunk gradients add up exactly.

bf16 mode: the fine depths come from the (bf16) coarse weights, so they differ slightly from the
oracle's; gradients are therefore also compared with the oracle evaluated AT the CUDA path's own fine
depths (`z_fine_override`), which isolates arithmetic error from sample placement."""
import os

import pytest
import torch

from oracle import synth
from oracle import upnerf_oracle as O
from oracle.make_golden import NET_CASES

pytestmark = pytest.mark.gpu
R, S, NI, N_IMG, M_SCHED, PROG = 4096, 64, 64, 763, 0.5, 0.3
CHUNK = 512


def _oracle(sds, cfgs, emb_w, rays0, img_idx, rng, cots, dtype, z_fine=None, operands=None):
    """Chunked oracle forward + backward in `dtype`; returns outputs (fp32) and gradients (fp64).
    `operands=torch.bfloat16`: dense-layer operands rounded to bf16 (upnerf_oracle.operand_rounding)."""
    with O.operand_rounding(operands):
        return _oracle_run(sds, cfgs, emb_w, rays0, img_idx, rng, cots, dtype, z_fine)


def _oracle_run(sds, cfgs, emb_w, rays0, img_idx, rng, cots, dtype, z_fine):
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    sd_o = {k: {n: cast(v).clone().requires_grad_(n != "progress") for n, v in sd.items()} for k, sd in sds.items()}
    emb_o = {k: cast(v).clone().requires_grad_(True) for k, v in emb_w.items()}
    rays_o = cast(rays0).clone().requires_grad_(True)
    outs = {}
    for i in range(0, R, CHUNK):
        sl = slice(i, i + CHUNK)
        r = O.RenderRng(perturb_rand=cast(rng["perturb_rand"][sl]), u=[cast(u[sl]) for u in rng["u"]])
        res = O.render_rays(sd_o, cfgs, emb_o, rays_o[sl], img_idx[sl], M_SCHED, PROG, N_samples=S, perturb=1.0,
                            N_importance=NI, rng=r, z_fine_override=None if z_fine is None else z_fine[sl])
        sum((res[k] * cast(cots[k][sl])).sum() for k in res if "weights" not in k).backward()
        for k, v in res.items():
            outs.setdefault(k, []).append(v.detach().float())
    outs = {k: torch.cat(v, 0) for k, v in outs.items()}
    grads = {"rays": rays_o.grad[:, :6].double()}
    for k, v in emb_o.items():
        if v.grad is not None:
            grads[f"emb_{k}"] = v.grad.double()
    for mk, sd in sd_o.items():
        for n, v in sd.items():
            if n != "progress" and v.grad is not None and float(v.grad.abs().max()) > 0:
                grads[f"{mk}.{n}"] = v.grad.double()
    return outs, grads


def _rel(a, b):
    return float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))


def _whole(grads, ref, keys):
    num = sum(float((grads[k].double().cpu() - ref[k]).pow(2).sum()) for k in keys)
    den = sum(float(ref[k].pow(2).sum()) for k in keys)
    return (num / den) ** 0.5


@pytest.fixture(scope="module")
def case():
    kw, _ = NET_CASES["full"]
    cfgs = {"nerf_coarse": O.NerfConfig(typ="coarse", **kw), "nerf_fine": O.NerfConfig(typ="fine", **kw)}
    sds = {"nerf_coarse": synth.nerf_state(cfgs["nerf_coarse"], 21, progress=PROG),
           "nerf_fine": synth.nerf_state(cfgs["nerf_fine"], 22, progress=PROG)}
    emb_w = synth.embeddings(N_IMG, cfgs["nerf_coarse"], 9)
    b = synth.ray_batch(R, N_IMG, 33)
    o, d = O.get_rays(b["directions"], b["c2w"])
    rays0 = torch.cat([o, d, b["ray_infos"]], 1)
    ns = round(M_SCHED * NI)
    rng = dict(perturb_rand=synth.uniform((R, S), 41, 0, 1),
               u=[synth.uniform((R, NI - ns), 42, 0, 1), synth.uniform((R, ns), 43, 0, 1)])
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    # cotangents: per-ray outputs weighted like a mean over rays and channels
    shapes = {"c_depth": (R,), "s_depth": (R,), "t_weight": (R,), "feat": (R, 384), "s_rgb": (R, 3),
              "c_weights": None, "s_weights": None}
    cots = {}
    for i, key in enumerate(sorted(shapes)):
        for j, typ in enumerate(("coarse", "fine")):
            if shapes[key] is None:
                continue
            c = synth.uniform(shapes[key], 500 + 2 * i + j)
            cots[f"{key}_{typ}"] = c / c[0].numel()
    out64, g64 = _oracle(sds, cfgs, emb_w, rays0, b["img_idx"], rng, cots, torch.float64)
    return dict(cfgs=cfgs, sds=sds, emb_w=emb_w, b=b, rays0=rays0, rng=rng, cots=cots, out64=out64, g64=g64)


def _cuda_run(case, precision, dev):
    from upnerf_b200.models.nerf import NeRF
    from upnerf_b200.models.rendering import render_rays

    models, embs = {}, {}
    for name, cfg in case["cfgs"].items():
        m = NeRF(cfg.typ, encode_feat=True, feat_dim=cfg.feat_dim, xyz_L=cfg.xyz_L, dir_L=cfg.dir_L,
                 appearance_dim=cfg.appearance_dim, candidate_dim=cfg.candidate_dim, c2f=cfg.c2f)
        m.load_state_dict(case["sds"][name])
        models[name] = m.to(dev)
    for k, w in case["emb_w"].items():
        e = torch.nn.Embedding(*w.shape)
        e.weight.data.copy_(w)
        embs[k] = e.to(dev)
    rays = case["rays0"].to(dev).requires_grad_(True)
    depths = {}
    res = render_rays(models=models, embeddings=embs, rays=rays, img_idx=case["b"]["img_idx"].to(dev),
                      sched_mult=M_SCHED, N_samples=S, perturb=1.0, N_importance=NI, encode_feat=True,
                      rng=case["rng"], precision=precision, return_depths=depths)
    sum((res[k] * case["cots"][k].to(dev)).sum() for k in res if "weights" not in k).backward()
    grads = {"rays": rays.grad[:, :6]}
    for k, e in embs.items():
        if e.weight.grad is not None:
            grads[f"emb_{k}"] = e.weight.grad
    for mk, mod in models.items():
        for pn, p in mod.named_parameters():
            if p.grad is not None:
                grads[f"{mk}.{pn}"] = p.grad
    return {k: v.detach().cpu() for k, v in res.items()}, grads, depths["z_fine"].cpu()


def _report(tag, grads, ref):
    keys = [k for k in ref if k != "rays" and not k.startswith("emb_")]
    per = {k: _rel(grads[k], ref[k]) for k in ref}
    whole = _whole(grads, ref, keys)
    worst = max(keys, key=lambda k: per[k])
    print(f"[{tag}] whole-network {whole:.3e}  worst tensor {worst} {per[worst]:.3e}  rays {per['rays']:.3e}  "
          + "  ".join(f"{k} {per[k]:.2e}" for k in ref if k.startswith("emb_")))
    for k in sorted(keys, key=lambda k: -per[k])[:6]:
        print(f"    {k:50s} {per[k]:.3e}")
    return whole, per, keys


def test_fp32_mode_at_baseline_size(cuda_dev, case):
    res, grads, z_fine = _cuda_run(case, "fp32", cuda_dev)
    out32, g32 = _oracle(case["sds"], case["cfgs"], case["emb_w"], case["rays0"], case["b"]["img_idx"], case["rng"],
                         case["cots"], torch.float32)
    assert list(res) == list(out32)
    for k, v in out32.items():
        err = float((res[k] - v).abs().max())
        if v.dim() == 2 and v.shape[1] == S + NI:
            # per-sample fine weights sit at resampled depths: a 1-ulp CDF difference moves a sample inside a
            # near-empty bin (test_kernels_gpu.test_sample_pdf_golden); bound the bulk, not the outliers
            assert float((res[k] - v).abs().mean()) <= 1e-5, (k, err)
            continue
        assert err <= 1e-4 * max(1.0, float(v.abs().max())), (k, err)
    # how far the reference's own fp32 arithmetic is from fp64 -- the noise floor of any fp32 comparison
    _report("oracle fp32 vs fp64", {k: v.float() for k, v in g32.items()}, case["g64"])
    whole, per, keys = _report("cuda fp32 vs oracle fp32", grads, g32)
    assert set(grads) >= set(g32)
    assert whole <= 1e-2, whole
    for k in keys:
        assert per[k] <= 1e-2, (k, per[k])
    for k in g32:
        if k.startswith("emb_"):
            assert per[k] <= 1e-2, (k, per[k])
    assert per["rays"] <= 1e-2, per["rays"]


def test_bf16_mode_at_baseline_size(cuda_dev, case):
    """north_star asks for parameter / pose gradients within 1e-2 relative.  fp32 mode meets it (above).
    In bf16 mode the bound is set by the data type, not by the kernels: rounding the dense layers'
    operands to bf16 flips the sign of the ~0.3 % of pre-activations that lie within rounding distance of
    zero, each flip is a 100 % error of one element of dY, and the flips accumulate down the ten-layer
    backward chain (error ~ sqrt(fraction flipped): ~1.6e-2 on the whole network, ~0.1 on layer 1).  The
    test proves that with the reference's own arithmetic: the oracle run with bf16-ROUNDED OPERANDS and
    exact (fp64) accumulation shows the same error against the exact fp64 oracle as the CUDA path does."""
    res, grads, z_fine = _cuda_run(case, "bf16", cuda_dev)
    for k, v in case["out64"].items():
        if v.dim() == 2 and v.shape[1] in (S, S + NI):
            continue
        err = float((res[k] - v).abs().max())
        assert err <= 2e-2 * max(1.0, float(v.abs().max())), (k, err)
    _report("cuda bf16 vs oracle fp64 (end to end, own fine depths differ)", grads, case["g64"])
    # the same comparison with the oracle evaluated at the CUDA path's fine depths
    args = (case["sds"], case["cfgs"], case["emb_w"], case["rays0"], case["b"]["img_idx"], case["rng"], case["cots"])
    _, g_at = _oracle(*args, torch.float64, z_fine=z_fine.double())
    whole, per, keys = _report("cuda bf16 vs oracle fp64 at the same fine depths", grads, g_at)
    # the reference arithmetic with bf16 operands (fp64 accumulation), same depths, against the same truth
    _, g_emu = _oracle(*args, torch.float64, z_fine=z_fine.double(), operands=torch.bfloat16)
    whole_e, per_e, _ = _report("oracle with bf16-rounded operands vs oracle fp64 (the data type's own error)",
                                {k: v.float() for k, v in g_emu.items()}, g_at)
    # 1. the CUDA path is no worse than the data type allows (25 % slack: flips are chaotic)
    assert whole <= 1.25 * whole_e + 1e-3, (whole, whole_e)
    for k in keys:
        assert per[k] <= 1.5 * per_e[k] + 5e-3, (k, per[k], per_e[k])
    assert per["rays"] <= 1.5 * per_e["rays"] + 5e-3, (per["rays"], per_e["rays"])
    # 2. absolute bounds achieved at the benchmark size (whole network norm-wise, per tensor, embeddings, pose path)
    assert whole <= 2.5e-2, whole
    for k in keys:
        assert per[k] <= 0.15, (k, per[k])
    for k in g_at:
        if k.startswith("emb_"):
            assert per[k] <= 2e-2, (k, per[k])
    assert per["rays"] <= 0.15, per["rays"]
