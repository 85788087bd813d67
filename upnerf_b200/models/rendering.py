"""`render_rays` / `sample_pdf` with the reference's signatures (models/rendering.py:7-50, 53-314).

`render_rays(models, embeddings, rays, img_idx, sched_mult, ...)` returns the same
phase-dependent dict of tensors as the reference and is differentiable with respect to the
rays (pose gradient), every NeRF parameter and the appearance / candidate embedding tables.
All arithmetic happens in libupnerf_b200.so (`upnerf_render_fwd/bwd`); there is no PyTorch
fallback -- CPU tensors raise.

Extra keyword arguments (ours): `precision="bf16"|"fp32"` (default: $UPNERF_PRECISION or
"bf16") and `rng=dict(perturb_rand=Tensor[R,S], u=[Tensor[R,n0], Tensor[R,n1]])` to inject the
uniforms the reference would draw for parity tests; `grad_sink=dict(coarse=, fine=,
coarse_a=, ...)` names fp32 buffers the backward ACCUMULATES parameter / embedding gradients into
(slices of a flat .grad buffer) instead of returning fresh tensors to autograd;
`after_fine_bwd=callable` is called between the fine and the coarse network's backward (gradient all-reduce
overlap); `return_depths=dict()` is filled with the sample depths `z_coarse` [R,S] / `z_fine` [R,S+N_importance].
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from .nerf import flat_parameters, net_config

__all__ = ["render_rays"]

_OUT_KEYS = ("c_weights", "s_weights", "c_depth", "s_depth", "t_weight", "feat", "s_rgb")


def default_precision() -> str:
    return os.environ.get("UPNERF_PRECISION", "bf16")


def _dtype_code(precision: str) -> int:
    if precision in ("bf16", "bfloat16"):
        return L.BF16
    if precision in ("fp32", "float32", "f32"):
        return L.F32
    raise L.UpnerfError(f"unknown precision {precision!r}")


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None):
    """Inverse-CDF sampling (models/rendering.py:7-50). `u` optionally injects the uniforms."""
    R = weights.shape[0]
    bins = bins.contiguous().float()
    weights = weights.detach().float()
    if weights.stride(1) != 1:
        weights = weights.contiguous()
    if u is None and not det:
        u = torch.rand(R, N_importance, device=bins.device)
    if u is not None:
        u = u.contiguous().float()
    out = torch.empty(R, N_importance, device=bins.device, dtype=torch.float32)
    L.sample_pdf(bins, weights, u, N_importance, eps, out)
    return out


def _phase_keys(cfg, m):
    cand = m < 1 and cfg.encode_candidate and cfg.candidate_dim > 0 and cfg.encode_feat
    keys = []
    if m < 1:
        if cand:
            keys += ["c_weights", "c_depth", "feat", "t_weight"]
        elif cfg.encode_feat:
            keys += ["s_weights", "feat"]
        else:
            raise NotImplementedError  # models/rendering.py:149-150
    if (m > 0 or not cfg.encode_feat) and "s_weights" not in keys:
        keys.append("s_weights")
    if m > 0 or not cfg.encode_feat:
        keys.append("s_rgb")
    keys.append("s_depth")
    return keys


def _shape(key, R, S, F):
    return {"c_weights": (R, S), "s_weights": (R, S), "feat": (R, F), "s_rgb": (R, 3)}.get(key, (R,))


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, rays, flat_c, flat_f, emb_ca, emb_fa, emb_cc, emb_fc):
        cfg, R, S, NI = meta["cfg"], rays.shape[0], meta["N_samples"], meta["N_importance"]
        dev = rays.device
        # outputs the loss never read arrive in backward as None (not as zero tensors): a pass without
        # any gradient is skipped by upnerf_render_bwd, unused [R,S] weight cotangents are never built
        ctx.set_materialize_grads(False)
        a = L.RenderArgs()
        a.cfg = cfg
        a.dtype = meta["dtype"]
        a.n_rays, a.n_samples, a.n_importance = R, S, NI
        a.n_importance_static = meta["n_static"]
        a.n_images = meta["n_images"]
        a.sched_mult = float(meta["sched_mult"])
        a.use_disp = int(meta["use_disp"])
        a.perturb = float(meta["perturb"])
        a.no_grad = int(meta["no_grad"])
        keep = [rays, meta["img_idx"], flat_c, flat_f, emb_ca, emb_fa, emb_cc, emb_fc,
                meta["perturb_rand"], meta["u0"], meta["u1"]]
        a.rays, a.img_idx = L._vp(rays), L._vp(meta["img_idx"])
        a.perturb_rand, a.u0, a.u1 = L._vp(meta["perturb_rand"]), L._vp(meta["u0"]), L._vp(meta["u1"])
        outs, names = [], []
        for which, flat, ea, ec, Sx in (("coarse", flat_c, emb_ca, emb_cc, S), ("fine", flat_f, emb_fa, emb_fc, S + NI)):
            if which == "fine" and NI == 0:
                continue
            io = getattr(a, which)
            io.params, io.emb_a, io.emb_c = L._vp(flat), L._vp(ea), L._vp(ec)
            for key in meta["keys"]:
                t = torch.empty(_shape(key, R, Sx, cfg.feat_dim), device=dev, dtype=torch.float32)
                setattr(io, key, t.data_ptr())
                outs.append(t)
                names.append(f"{key}_{which}")
        if meta.get("depths") is not None:        # optional debug outputs: the sample depths of both passes
            zc = torch.empty(R, S, device=dev, dtype=torch.float32)
            zf = torch.empty(R, S + NI, device=dev, dtype=torch.float32) if NI > 0 else None
            a.z_coarse, a.z_fine = zc.data_ptr(), L._vp(zf)
            meta["depths"].update(z_coarse=zc, z_fine=zf)
        a.workspace_bytes = 0
        nbytes = L.render_workspace_bytes(a)
        # inference chunk loops pass `workspace_cache` (a dict): every chunk then runs in the SAME workspace, which
        # lets the chunks after the first skip the parameter-only prologue (`reuse_packed`)
        wsc = meta.get("workspace_cache")
        ws = None if wsc is None else wsc.get(nbytes)
        a.reuse_packed = 0
        if ws is None or ws.device != dev:
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            if wsc is not None:
                wsc[nbytes] = ws
        elif meta.get("reuse_packed") and a.no_grad:
            a.reuse_packed = 1
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        L.render_fwd(a)
        ctx.args, ctx.keep, ctx.ws, ctx.names, ctx.meta = a, keep, ws, names, meta
        ctx.shapes = (flat_c.shape, None if flat_f is None else flat_f.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        a, meta = ctx.args, ctx.meta
        rays, _, flat_c, flat_f, emb_ca, emb_fa, emb_cc, emb_fc = ctx.keep[:8]
        dev = rays.device
        hold = []
        cfg = meta["cfg"]
        feat_phase = "feat" in meta["keys"]
        for name, g in zip(ctx.names, gouts):
            key, which = name.rsplit("_", 1)
            io = getattr(a, which)
            if g is None:
                if key == "feat" and feat_phase:
                    g = torch.zeros(rays.shape[0], cfg.feat_dim, device=dev)
                else:
                    setattr(io, "g_" + key, None)
                    continue
            g = g.contiguous().float()
            hold.append(g)
            setattr(io, "g_" + key, g.data_ptr())
        need = ctx.needs_input_grad
        d_rays = torch.zeros_like(rays) if need[1] else None
        a.d_rays = L._vp(d_rays)
        grads = {}
        sink = meta.get("grad_sink") or {}

        def dest(key, like):
            """Gradient destination: the caller's accumulation buffer (`grad_sink`, e.g. a slice of
            a flat .grad buffer -- the kernels ACCUMULATE, so autograd gets None for that input) or
            a fresh zero tensor handed back to autograd."""
            if like is None:
                return None
            t = sink.get(key)
            if t is not None:
                if t.numel() != like.numel() or t.dtype != torch.float32 or not t.is_contiguous():
                    raise L.UpnerfError(f"render_rays: grad_sink[{key!r}] does not match its parameter")
                hold.append(t)
                return t
            grads[key] = torch.zeros_like(like)
            return grads[key]

        # needs_input_grad order = forward's arguments: (meta, rays, flat_c, flat_f, emb_ca, emb_fa, emb_cc, emb_fc)
        wants = {"coarse": need[2], "fine": need[3], "coarse_a": need[4], "fine_a": need[5], "coarse_c": need[6],
                 "fine_c": need[7]}
        for which, flat, ea, ec in (("coarse", flat_c, emb_ca, emb_cc), ("fine", flat_f, emb_fa, emb_fc)):
            if flat is None:
                continue
            io = getattr(a, which)
            # a frozen network (test-time optimisation): NULL d_params skips every weight-gradient launch
            frozen = not wants[which] and sink.get(which) is None
            io.d_params = None if frozen else dest(which, flat).data_ptr()
            ga = None if (not wants[which + "_a"] and sink.get(which + "_a") is None) else dest(which + "_a", ea)
            gc = None if (not wants[which + "_c"] and sink.get(which + "_c") is None) else dest(which + "_c", ec)
            io.d_emb_a, io.d_emb_c = L._vp(ga), L._vp(gc)
        hook = meta.get("after_fine_bwd")
        if hook is not None and a.n_importance > 0:
            # data-parallel training: the fine network's gradients are final after its pass -- the caller's
            # hook starts their all-reduce, which then runs under the coarse pass
            L.render_bwd_passes(a, 1)
            hook()
            L.render_bwd_passes(a, 2)
        else:
            L.render_bwd(a)
        return (None, d_rays, grads.get("coarse"), grads.get("fine"), grads.get("coarse_a"),
                grads.get("fine_a"), grads.get("coarse_c"), grads.get("fine_c"))


def render_rays(models, embeddings, rays, img_idx, sched_mult, N_samples=64, use_disp=False, perturb=0,
                N_importance=0, test_time=False, encode_feat=True, **kwargs):
    """Drop-in for the reference `render_rays` (models/rendering.py:53-314)."""
    if not rays.is_cuda:
        raise L.UpnerfError("render_rays: upnerf_b200 runs on CUDA tensors only (no CPU fallback)")
    coarse = models["nerf_coarse"]
    fine = models["nerf_fine"] if N_importance > 0 else None
    cfg = net_config(coarse)
    if bool(coarse.encode_feat) != bool(encode_feat):
        raise L.UpnerfError("render_rays: encode_feat disagrees with the model")
    R = rays.shape[0]
    dev = rays.device
    m = sched_mult
    rng = kwargs.get("rng") or {}
    perturb_rand = u0 = u1 = None
    n_static = 0
    cand_fine = fine is not None and fine.encode_candidate
    if perturb > 0:
        perturb_rand = rng.get("perturb_rand")
        if perturb_rand is None:
            perturb_rand = torch.rand(R, N_samples, device=dev)
    if N_importance > 0:
        two = cand_fine and 0 < m < 1
        if two:
            n_static = round(m * N_importance)          # Python banker's rounding, as the reference
        if perturb > 0:
            us = list(rng.get("u") or [])
            n0 = N_importance - n_static
            u0 = us[0] if us else torch.rand(R, n0, device=dev)
            if two:
                u1 = us[1] if len(us) > 1 else torch.rand(R, n_static, device=dev)
    as_f32 = lambda t: None if t is None else t.to(dev).contiguous().float()
    meta = dict(cfg=cfg, N_samples=N_samples, N_importance=N_importance, n_static=n_static,
                sched_mult=m, use_disp=use_disp, perturb=perturb, keys=_phase_keys(cfg, m),
                img_idx=img_idx.contiguous().long(), perturb_rand=as_f32(perturb_rand), u0=as_f32(u0),
                u1=as_f32(u1), dtype=_dtype_code(kwargs.get("precision") or default_precision()),
                n_images=0, no_grad=False, grad_sink=kwargs.get("grad_sink"),
                depths=kwargs.get("return_depths"), after_fine_bwd=kwargs.get("after_fine_bwd"),
                workspace_cache=kwargs.get("workspace_cache"), reuse_packed=bool(kwargs.get("reuse_packed")))
    emb = lambda k: embeddings[k].weight if k in embeddings else None
    ea_c = emb("coarse_a") if coarse.encode_appearance else None
    ec_c = emb("coarse_c") if coarse.encode_candidate else None
    ea_f = emb("fine_a") if fine is not None and fine.encode_appearance else None
    ec_f = emb("fine_c") if fine is not None and fine.encode_candidate else None
    for e in (ea_c, ec_c, ea_f, ec_f):
        if e is not None:
            meta["n_images"] = e.shape[0]
    flat_c = flat_parameters(coarse)
    flat_f = flat_parameters(fine) if fine is not None else None
    # inference (the reference's validation / tto render runs under no_grad): nothing is kept for backward
    meta["no_grad"] = not (torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (rays, flat_c, flat_f, ea_c, ea_f, ec_c, ec_f)))
    outs = _RenderFn.apply(meta, rays.contiguous().float(), flat_c, flat_f, ea_c, ea_f, ec_c, ec_f)
    results, it = {}, iter(outs)
    for which in ("coarse", "fine") if fine is not None else ("coarse",):
        typ = (coarse if which == "coarse" else fine).typ
        for key in meta["keys"]:
            results[f"{key}_{typ}"] = next(it)
    return results
