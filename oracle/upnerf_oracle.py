"""CPU oracle for the UP-NeRF train/render hot path.

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch restatement, in plain PyTorch
tensor ops on the CPU (fp32 by default, fp64 on request), of the arithmetic the reference
(mlvlab/UP-NeRF, /root/reference) performs on the path this repo accelerates.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product package `upnerf_b200` never does.

Pinning: every function here is checked against outputs of the *real* reference code
(imported from /root/reference in the build container by `oracle/make_golden.py`) that are
committed under `tests/golden/` -- see `tests/test_oracle_golden.py`.  The reference has no
tests or golden vectors of its own, so these generated fixtures are
what pins parity.

Each function cites the reference file:line it restates.  Networks are passed as plain
`dict[str, Tensor]` using the reference's `state_dict()` key names, so the same dict
drives the reference module, this oracle and the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# SE(3) pose refinement and ray generation
# --------------------------------------------------------------------------------------


def _taylor_series(theta: torch.Tensor, kind: str, order: int = 10) -> torch.Tensor:
    """Order-10 Taylor sums used instead of closed forms (utils/camera.py:126-152).

    kind "A": sin(t)/t, "B": (1-cos t)/t^2, "C": (t-sin t)/t^3.  Terms are added one by one
    in the input dtype with a Python-float factorial denominator, as the reference does.
    """
    total = torch.zeros_like(theta)
    denom = 1.0
    for i in range(order + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        elif kind == "C":
            denom *= (2 * i + 2) * (2 * i + 3)
        else:  # pragma: no cover
            raise ValueError(kind)
        total = total + (-1) ** i * theta ** (2 * i) / denom
    return total


def skew(w: torch.Tensor) -> torch.Tensor:
    """[w]_x cross-product matrix (utils/camera.py:113-124)."""
    wx, wy, wz = w.unbind(-1)
    zero = torch.zeros_like(wx)
    rows = [torch.stack(r, -1) for r in ((zero, -wz, wy), (wz, zero, -wx), (-wy, wx, zero))]
    return torch.stack(rows, -2)


def se3_exp(wu: torch.Tensor) -> torch.Tensor:
    """Exponential map se(3) -> [R|t] of shape (...,3,4) (utils/camera.py:87-98)."""
    w, u = wu[..., :3], wu[..., 3:]
    K = skew(w)
    theta = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, dtype=wu.dtype, device=wu.device)
    A, B, C = (_taylor_series(theta, k) for k in "ABC")
    K2 = K @ K
    R = eye + A * K + B * K2
    V = eye + B * K + C * K2
    return torch.cat([R, V @ u[..., None]], -1)


def compose_pair(first: torch.Tensor, second: torch.Tensor) -> torch.Tensor:
    """Pose `second o first` on [R|t] matrices (utils/camera.py:51-58)."""
    Ra, ta = first[..., :3], first[..., 3:]
    Rb, tb = second[..., :3], second[..., 3:]
    return torch.cat([Rb @ Ra, Rb @ ta + tb], -1)


def compose(poses) -> torch.Tensor:
    """Left-to-right composition of a list of poses (utils/camera.py:43-49)."""
    out = poses[0]
    for p in poses[1:]:
        out = compose_pair(out, p)
    return out


def get_rays(directions: torch.Tensor, c2w: torch.Tensor):
    """World-space origins and unit directions (utils/ray.py:30-67).

    Batched branch (:44-56): one pose per ray.  Otherwise (:57-65) a single (3,4) pose.
    """
    if c2w.dim() == 3 and directions.dim() == 2 and c2w.shape[0] == directions.shape[0]:
        d = torch.einsum("rij,rj->ri", c2w[:, :, :3], directions)
        d = d / d.norm(dim=-1, keepdim=True)
        o = c2w[..., 3]
        return o.reshape(-1, 3), d.reshape(-1, 3)
    d = directions @ c2w[:, :3].T
    d = d / d.norm(dim=-1, keepdim=True)
    o = c2w[:, 3].expand(d.shape)
    return o.reshape(-1, 3), d.reshape(-1, 3)


def refine_and_cast(se3_table: torch.Tensor, img_idx: torch.Tensor, c2w: torch.Tensor,
                    directions: torch.Tensor):
    """Pose refinement + ray casting of a training batch (models/nerf_system.py:158-166)."""
    refined = compose([se3_exp(se3_table[img_idx]), c2w])
    return get_rays(directions, refined)


# --------------------------------------------------------------------------------------
# Positional encoding and the NeRF MLP
# --------------------------------------------------------------------------------------


@dataclass
class NerfConfig:
    """Constructor arguments of the reference NeRF (models/nerf.py:6-19)."""

    typ: str = "coarse"
    D: int = 8
    W: int = 256
    skips: tuple = (4,)
    encode_feat: bool = True
    feat_dim: int = 384
    xyz_L: int = 10
    dir_L: int = 4
    appearance_dim: int = 48
    candidate_dim: int = 16
    c2f: tuple | None = (0.1, 0.5)

    @property
    def encode_appearance(self) -> bool:
        return self.appearance_dim > 0

    @property
    def encode_candidate(self) -> bool:
        return self.candidate_dim > 0


def c2f_weights(L: int, progress: float, c2f, dtype=torch.float32, device=None) -> torch.Tensor:
    """Per-band coarse-to-fine weights (models/nerf.py:137-142); ones when c2f is None."""
    if c2f is None:
        return torch.ones(L, dtype=dtype, device=device)
    start, end = c2f
    alpha = (torch.tensor(progress, dtype=dtype, device=device) - start) / (end - start) * L
    k = torch.arange(L, dtype=dtype, device=device)
    return (1 - ((alpha - k).clamp(0, 1) * math.pi).cos()) / 2


def positional_encoding(x: torch.Tensor, L: int, progress: float = 1.0, c2f=None) -> torch.Tensor:
    """BARF-style encoding (models/nerf.py:126-147).

    Output layout per row: [x, then for each coordinate: w_k sin(x f_k) (k<L), w_k cos(x f_k)
    (k<L)], f_k = 2^k * pi computed in fp32.
    """
    freq = (2 ** torch.arange(L, dtype=torch.float32, device=x.device) * math.pi).to(x.dtype)
    spec = x[..., None] * freq
    w = c2f_weights(L, progress, c2f, x.dtype, x.device)
    enc = torch.stack([spec.sin() * w, spec.cos() * w], -2)  # (..., 3, 2, L)
    return torch.cat([x, enc.reshape(*x.shape[:-1], -1)], -1)


class _RoundOperand(torch.autograd.Function):
    """x -> x rounded to `dtype` (value kept in x's own dtype); the incoming gradient is rounded the same
    way.  Emulates a path that STORES activations / their gradients in a narrow type and accumulates wide."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


# None: exact arithmetic in the tensors' own dtype (the reference).  torch.bfloat16: every dense layer of
# NeRF.forward sees its input activations and its weight matrix rounded to bf16 and accumulates in the
# tensors' dtype -- "the reference's arithmetic with bf16 operands", the floor of ANY bf16 tensor-core
# implementation.  Used by tests/test_baseline_size_gpu.py to show where the CUDA path's bf16 gradient
# error comes from (ReLU sign flips of near-zero pre-activations), not by the parity oracle itself.
OPERAND_ROUNDING = None


class operand_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_ROUNDING
        self.prev, OPERAND_ROUNDING = OPERAND_ROUNDING, self.dtype

    def __exit__(self, *exc):
        global OPERAND_ROUNDING
        OPERAND_ROUNDING = self.prev


def _linear(p: dict, name: str, x: torch.Tensor) -> torch.Tensor:
    w = p[name + ".weight"]
    if OPERAND_ROUNDING is not None and w.shape[0] > 3:        # the N = 1 / 3 heads stay fp32 row-dots
        x, w = _RoundOperand.apply(x, OPERAND_ROUNDING), _RoundOperand.apply(w, OPERAND_ROUNDING)
    return F.linear(x, w, p[name + ".bias"])


def nerf_forward(p: dict, cfg: NerfConfig, xyz: torch.Tensor, dirs: torch.Tensor,
                 a_emb: torch.Tensor | None, c_emb: torch.Tensor | None, sched_mult: float,
                 progress: float) -> dict:
    """NeRF.forward on flat (M,.) inputs (models/nerf.py:80-124).

    `p` uses the reference state_dict names ("xyz_encoding_1.0.weight", ...).
    """
    out = {}
    pe = positional_encoding(xyz, cfg.xyz_L, progress, cfg.c2f)
    h = pe
    for i in range(cfg.D):
        if i in cfg.skips:
            h = torch.cat([pe, h], 1)
        h = F.relu(_linear(p, f"xyz_encoding_{i + 1}.0", h))
    out["s_sigma"] = F.softplus(_linear(p, "share_sigma.0", h))
    hf = _linear(p, "xyz_encoding_final", h)

    def rgb_head(front):
        parts = [front, positional_encoding(dirs, cfg.dir_L, progress, cfg.c2f)]
        if cfg.encode_appearance:
            parts.append(a_emb)
        q = F.relu(_linear(p, "rgb_share_layer.0", torch.cat(parts, 1)))
        return torch.sigmoid(_linear(p, "rgb_share_layer.2", q))

    def candidate_trunk():
        g = F.relu(_linear(p, "candidate_encoding.0", torch.cat([hf, c_emb], 1)))
        g = F.relu(_linear(p, "candidate_encoding.2", g))
        out["c_sigma"] = F.softplus(_linear(p, "candidate_sigma.0", g))
        return g

    if cfg.encode_feat:
        out["s_feat"] = _linear(p, "feat_share_layer", hf)
        if sched_mult < 1 and cfg.encode_candidate:
            out["c_feat"] = _linear(p, "feat_candidate_layer", candidate_trunk())
        if sched_mult > 0:
            out["s_rgb"] = rgb_head(out["s_feat"])
    else:
        out["s_rgb"] = rgb_head(hf)
        if sched_mult < 1:
            out["c_rgb"] = _linear(p, "rgb_candidate_layer", candidate_trunk())
    return out


# --------------------------------------------------------------------------------------
# Volume rendering: compositing, hierarchical sampling, render_rays
# --------------------------------------------------------------------------------------


def _exclusive_cumprod(x: torch.Tensor) -> torch.Tensor:
    ones = torch.ones_like(x[:, :1])
    return torch.cumprod(torch.cat([ones, x], -1)[:, :-1], -1)


def composite(results: dict, typ: str, net_out: dict, z: torch.Tensor, sched_mult: float,
              encode_candidate: bool, encode_feat: bool) -> None:
    """Alpha compositing of one network pass (models/rendering.py:124-219).

    `net_out` holds per-sample (R,S[,C]) tensors.  Fills `results` with the phase-dependent
    keys of SURVEY.md section 3.2.
    """
    delta = torch.cat([z[:, 1:] - z[:, :-1], 1e2 * torch.ones_like(z[:, :1])], -1)
    s_alpha = 1 - torch.exp(-delta * net_out["s_sigma"])
    wsum = lambda w, v: (w[..., None] * v).sum(1)

    if sched_mult < 1:
        if not encode_candidate:
            if not encode_feat:
                raise NotImplementedError  # rendering.py:149-150
            w = s_alpha * _exclusive_cumprod(1 - s_alpha)
            results[f"s_weights_{typ}"] = w
            results[f"feat_{typ}"] = wsum(w, net_out["s_feat"])
        else:
            c_alpha = 1 - torch.exp(-delta * net_out["c_sigma"])
            alpha = 1 - torch.exp(-delta * (net_out["s_sigma"] + net_out["c_sigma"]))
            T = _exclusive_cumprod(1 - alpha)
            s_w, c_w, w = s_alpha * T, c_alpha * T, alpha * T
            results[f"c_weights_{typ}"] = w
            results[f"c_depth_{typ}"] = (w * z).sum(1)
            if encode_feat:
                results[f"feat_{typ}"] = wsum(s_w, net_out["s_feat"]) + wsum(c_w, net_out["c_feat"])
            else:
                results[f"c_rgb_{typ}"] = wsum(s_w, net_out["s_rgb"]) + wsum(c_w, net_out["c_rgb"])
            results[f"t_weight_{typ}"] = c_w.sum(1)

    static_w = s_alpha * _exclusive_cumprod(1 - s_alpha)
    if sched_mult > 0:
        results[f"s_weights_{typ}"] = static_w
        results[f"s_rgb_{typ}"] = wsum(static_w, net_out["s_rgb"])
    results[f"s_depth_{typ}"] = (static_w * z).sum(1)


def pdf_cdf(weights: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Padded CDF of the (detached) coarse weights (models/rendering.py:19-23)."""
    w = weights + eps
    pdf = w / w.sum(-1, keepdim=True)
    return torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n: int, det: bool = False,
               eps: float = 1e-5, u: torch.Tensor | None = None, return_inds: bool = False):
    """Inverse-CDF sampling (models/rendering.py:7-50).

    bins (R,S-1) interval mid-points, weights (R,S-2).  `u` overrides the random draw
    (rendering.py:29) so that CPU and GPU runs share the same uniforms.
    """
    R, nw = weights.shape
    cdf = pdf_cdf(weights, eps)
    if u is None:
        u = (torch.linspace(0, 1, n, dtype=bins.dtype, device=bins.device).expand(R, n) if det
             else torch.rand(R, n, dtype=bins.dtype, device=bins.device))
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(nw)
    cdf_lo, cdf_hi = cdf.gather(1, below), cdf.gather(1, above)
    bin_lo, bin_hi = bins.gather(1, below), bins.gather(1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    samples = bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)
    return (samples, inds) if return_inds else samples


def stratified_z(near: torch.Tensor, far: torch.Tensor, S: int, use_disp: bool = False,
                 perturb: float = 0.0, perturb_rand: torch.Tensor | None = None) -> torch.Tensor:
    """Coarse sample depths (models/rendering.py:231-249). near/far are (R,1)."""
    s = torch.linspace(0, 1, S, dtype=near.dtype, device=near.device)
    z = 1 / (1 / near * (1 - s) + 1 / far * s) if use_disp else near * (1 - s) + far * s
    z = z.expand(near.shape[0], S)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        if perturb_rand is None:
            perturb_rand = torch.rand_like(z)
        z = lower + (upper - lower) * (perturb * perturb_rand)
    return z


@dataclass
class RenderRng:
    """Explicit random inputs in the order the reference draws them."""

    perturb_rand: torch.Tensor | None = None      # (R, S_c), rendering.py:248
    u: list = field(default_factory=list)          # one (R, n) per sample_pdf call, in call order


def render_rays(nets: dict, cfgs: dict, embeddings: dict, rays: torch.Tensor, img_idx: torch.Tensor,
                sched_mult: float, progress: float, N_samples: int = 64, use_disp: bool = False,
                perturb: float = 0.0, N_importance: int = 0, encode_feat: bool = True,
                rng: RenderRng | None = None, return_aux: bool = False,
                z_fine_override: torch.Tensor | None = None) -> dict:
    """Coarse + fine rendering of a ray batch (models/rendering.py:53-314).

    nets/cfgs: {"nerf_coarse": ..., "nerf_fine": ...}; embeddings: {"coarse_a": (N_img,48) ...}
    weight tensors.  `progress` is the NeRF.progress scalar (drives the c2f mask).
    """
    rng = rng or RenderRng()
    R = rays.shape[0]
    o, d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z = stratified_z(near, far, N_samples, use_disp, perturb, rng.perturb_rand)
    results, aux = {}, {}

    def run(which: str, z_vals: torch.Tensor):
        p, cfg = nets[f"nerf_{which}"], cfgs[f"nerf_{which}"]
        S = z_vals.shape[1]
        xyz = (o[:, None, :] + d[:, None, :] * z_vals[..., None]).reshape(-1, 3)
        dirs = d.detach()[:, None, :].expand(R, S, 3).reshape(-1, 3)          # rendering.py:104-106
        rep = lambda e: e[:, None, :].expand(R, S, e.shape[-1]).reshape(R * S, -1)
        a = rep(embeddings[f"{which}_a"][img_idx]) if cfg.encode_appearance else None
        c = rep(embeddings[f"{which}_c"][img_idx]) if cfg.encode_candidate else None
        out = nerf_forward(p, cfg, xyz, dirs, a, c, sched_mult, progress)
        out = {k: (v.reshape(R, S) if "sigma" in k else v.reshape(R, S, -1)) for k, v in out.items()}
        composite(results, cfg.typ, out, z_vals, sched_mult, cfg.encode_candidate, encode_feat)

    run("coarse", z)
    aux["z_coarse"] = z
    if N_importance > 0:
        cfg_f = cfgs["nerf_fine"]
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        det = perturb == 0
        us = list(rng.u)
        draw = lambda key, n: sample_pdf(mid, results[key][:, 1:-1].detach(), n, det=det,
                                         u=us.pop(0) if us else None)
        if cfg_f.encode_candidate and sched_mult == 0:                        # rendering.py:268-275
            new = [draw("c_weights_coarse", N_importance)]
        elif cfg_f.encode_candidate and 0 < sched_mult < 1:                   # rendering.py:276-290
            n_static = round(sched_mult * N_importance)
            cand = draw("c_weights_coarse", N_importance - n_static)
            stat = draw("s_weights_coarse", n_static)
            new = [stat, cand]
        else:                                                                 # rendering.py:291-307
            new = [draw("s_weights_coarse", N_importance)]
        z_fine = torch.sort(torch.cat([z, *new], -1), -1)[0]
        if z_fine_override is not None:       # tests: evaluate the fine pass at depths another path produced
            assert z_fine_override.shape == z_fine.shape
            z_fine = z_fine_override.to(z_fine.dtype)
        aux["z_fine"] = z_fine
        run("fine", z_fine)
    if return_aux:
        results["_aux"] = aux
    return results


# --------------------------------------------------------------------------------------
# Per-ray tail: TransientNet, loss, schedule (adjacent to the hot path)
# --------------------------------------------------------------------------------------


def transient_net(p: dict, feat: torch.Tensor, img_idx: torch.Tensor, beta_min: float = 0.1) -> dict:
    """2-D per-ray transient MLP (models/transient_net.py:27-38)."""
    h = feat
    for i in (0, 2, 4, 6):
        h = F.relu(_linear(p, f"feat_encoder.{i}", h))
    fin = _linear(p, "final_encoder", h)
    t = F.relu(_linear(p, "t_encoder.0", torch.cat([fin, p["embedding_t.weight"][img_idx]], -1)))
    alpha = torch.sigmoid(_linear(p, "alpha_layer.0", h))
    rgb = torch.sigmoid(_linear(p, "rgb_layer.0", t))
    beta = F.softplus(_linear(p, "beta_layer.0", t)) * alpha + beta_min
    return {"alpha": alpha, "rgb": rgb, "beta": beta}


def blend_transient(results: dict, t_out: dict) -> None:
    """rgb_coarse/rgb_fine/t_beta/t_alpha (models/nerf_system.py:128-144)."""
    a, c = t_out["alpha"], t_out["rgb"]
    results["rgb_coarse"] = results["s_rgb_coarse"] * (1 - a.detach()) + c.detach() * a.detach()
    results["rgb_fine"] = results["s_rgb_fine"] * (1 - a) + c * a
    results["t_beta"] = t_out["beta"]
    results["t_alpha"] = a


def upnerf_loss(res: dict, rgb_t: torch.Tensor, feat_t: torch.Tensor, depth_t: torch.Tensor, m: float,
                depth_mult: float = 1e-3, alpha_reg: float = 1.0, fine: bool = True) -> dict:
    """UPNeRFLoss with encode_feat=True (losses.py:21-64)."""
    out = {}
    for tag, typ in (("c", "coarse"), ("f", "fine")):
        if typ == "fine" and not fine:
            break
        if m < 1:
            ld = (res[f"s_depth_{typ}"] - depth_t).abs()
            if f"t_weight_{typ}" in res:
                ld = ld * (1 - res[f"t_weight_{typ}"].detach())
            out[f"l_depth_{tag}"] = ld.mean() * depth_mult * (1 - m)
            out[f"l_feat_{tag}"] = ((res[f"feat_{typ}"] - feat_t) ** 2).mean() * (1 - m)
        if m > 0:
            sq = (res[f"s_rgb_{typ}"] - rgb_t) ** 2
            if typ == "coarse":
                out["l_rgb_c"] = sq.mean() * m / 2
            else:
                out["l_rgb_f"] = (sq / (2 * res["t_beta"] ** 2)).mean() * m
                out["l_beta"] = torch.log(res["t_beta"]).mean() * m
                out["l_alpha"] = res["t_alpha"].mean() * alpha_reg * m
    return out


def schedule_mult(progress: float, schedule=(0.1, 0.5)) -> float:
    """Candidate-head schedule (models/nerf_system.py:452-461)."""
    s, e = schedule
    if progress < s:
        return 0
    if progress > e:
        return 1
    return (1 - math.cos(math.pi * (progress - s) / (e - s))) / 2


def predicted_depth(depth_scale: torch.Tensor, img_idx: torch.Tensor, inv_depth: torch.Tensor,
                    near: float, far: float) -> torch.Tensor:
    """Mono-depth affine correction (models/nerf_system.py:169-177)."""
    scale, shift = depth_scale[img_idx].unbind(1)
    inv = inv_depth * torch.exp(scale) + shift
    inv = torch.where(inv < 1 / far, torch.full_like(inv, 1 / far), inv)
    dep = 1.0 / inv
    return torch.where(dep < near, torch.full_like(dep, near), dep)


def psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """utils/metric.py:19-20."""
    return -10 * torch.log10(((pred - target) ** 2).mean())
