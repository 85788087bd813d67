"""ctypes binding of libupnerf_b200.so (the C ABI declared in include/upnerf_b200.h).

There is deliberately no fallback: if the shared library has not been built, or a call
returns a non-zero status, this module raises.  torch is used only for device memory and
streams -- tensors are handed to the library as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libupnerf_b200.so"
_lib = None


class UpnerfError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """Mirror of `upnerf_epilogue` (include/upnerf_b200.h)."""

    _fields_ = [
        ("bias", C.c_void_p),
        ("ray_bias", C.c_void_p),
        ("rows_per_ray", C.c_int),
        ("rank1_row", C.c_void_p),
        ("rank1_col", C.c_void_p),
        ("aux", C.c_void_p),
        ("ldaux", C.c_int64),
        ("aux_mode", C.c_int),
        ("act", C.c_int),
        ("n_heads", C.c_int),
        ("head_w", C.c_void_p),
        ("head_b", C.c_void_p),
        ("head_act", C.c_int),
        ("head_out", C.c_void_p),
        ("head_col_begin", C.c_int),
    ]


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise UpnerfError(
            f"{_LIB_PATH} is missing: build it with `python -m upnerf_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA path)"
        )
    _lib = C.CDLL(str(_LIB_PATH))
    _lib.upnerf_last_error.restype = C.c_char_p
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().upnerf_last_error().decode("utf-8", "replace")
        raise UpnerfError(f"{what} failed (status {status}): {msg}")


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise UpnerfError("upnerf_b200 kernels take CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    """Kernels the library has launched so far in this process (upnerf_launch_count)."""
    f = lib().upnerf_launch_count
    f.restype = C.c_longlong
    return int(f())


def launch_count_add(n: int) -> None:
    """Credit the kernel nodes of a replayed CUDA graph (captured from this library's launches)."""
    f = lib().upnerf_launch_count_add
    f.argtypes = [C.c_longlong]
    f.restype = None
    f(int(n))


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_epilogue(bias=None, ray_bias=None, rows_per_ray=0, rank1_row=None, rank1_col=None,
                  aux=None, ldaux=0, aux_mode=0, act=0, head_w=None, head_b=None, head_act=0,
                  head_out=None) -> Epilogue:
    ep = Epilogue()
    ep.bias = ptr(bias)
    ep.ray_bias = ptr(ray_bias)
    ep.rows_per_ray = int(rows_per_ray)
    ep.rank1_row = ptr(rank1_row)
    ep.rank1_col = ptr(rank1_col)
    ep.aux = ptr(aux)
    ep.ldaux = int(ldaux)
    ep.aux_mode = int(aux_mode)
    ep.act = int(act)
    ep.n_heads = 0 if head_w is None else int(head_w.shape[0])
    ep.head_w = ptr(head_w)
    ep.head_b = ptr(head_b)
    ep.head_act = int(head_act)
    ep.head_out = ptr(head_out)
    return ep


def _i64(x) -> C.c_int64:
    return C.c_int64(int(x))


def gemm_bf16(A, B, C_out, M, N, K, lda=None, ldb=None, ldc=None, ep: Epilogue | None = None):
    """C[M,N] = epi(A[M,K] @ B[N,K]^T) on tcgen05 (see upnerf_gemm_bf16)."""
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    ldc = C_out.stride(0) if ldc is None else ldc
    st = lib().upnerf_gemm_bf16(ptr(A), _i64(lda), ptr(B), _i64(ldb), ptr(C_out), _i64(ldc),
                                _i64(M), C.c_int(N), C.c_int(K),
                                C.byref(ep) if ep is not None else None, stream_ptr())
    check(st, "upnerf_gemm_bf16")


def gemm2_bf16(A1, A2, B, C_out, M, N, K1, K2, ep: Epilogue | None = None):
    """C[M,N] = epi([A1 | A2] @ B[N,K1+K2]^T) on tcgen05 (see upnerf_gemm2_bf16)."""
    st = lib().upnerf_gemm2_bf16(ptr(A1), _i64(A1.stride(0)), C.c_int(K1), ptr(A2), _i64(A2.stride(0)), C.c_int(K2),
                                 ptr(B), _i64(B.stride(0)), ptr(C_out), _i64(C_out.stride(0)), _i64(M), C.c_int(N),
                                 C.byref(ep) if ep is not None else None, stream_ptr())
    check(st, "upnerf_gemm2_bf16")


def wgrad_bf16(dY, X, dW, db, M, N, K, segs, lddy=None, ldx=None, lddw=None):
    """dW[n, map(k)] += dY^T X, db[n] += colsum(dY) on tcgen05 (see upnerf_wgrad_bf16)."""
    lddy = dY.stride(0) if lddy is None else lddy
    ldx = X.stride(0) if ldx is None else ldx
    lddw = dW.stride(0) if lddw is None else lddw
    n = len(segs)
    arr = C.c_int * n
    src = arr(*[s[0] for s in segs])
    ln = arr(*[s[1] for s in segs])
    dst = arr(*[s[2] for s in segs])
    st = lib().upnerf_wgrad_bf16(ptr(dY), _i64(lddy), ptr(X), _i64(ldx), ptr(dW), _i64(lddw),
                                 ptr(db), _i64(M), C.c_int(N), C.c_int(K), C.c_int(n), src, ln, dst,
                                 stream_ptr())
    check(st, "upnerf_wgrad_bf16")


def wgrad_bf16_det(dY, X, dW, db, M, N, K, segs, lddy=None, ldx=None, lddw=None):
    """Deterministic (atomic-free) variant of `wgrad_bf16` (see upnerf_wgrad_bf16_det)."""
    import torch

    lddy = dY.stride(0) if lddy is None else lddy
    ldx = X.stride(0) if ldx is None else ldx
    lddw = dW.stride(0) if lddw is None else lddw
    n = len(segs)
    arr = C.c_int * n
    src = arr(*[s[0] for s in segs])
    ln = arr(*[s[1] for s in segs])
    dst = arr(*[s[2] for s in segs])
    f = lib().upnerf_wgrad_det_workspace_bytes
    f.restype = C.c_uint64
    nbytes = int(f(C.c_int(N), C.c_int(K)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dY.device)
    st = lib().upnerf_wgrad_bf16_det(ptr(dY), _i64(lddy), ptr(X), _i64(ldx), ptr(dW), _i64(lddw),
                                     ptr(db), _i64(M), C.c_int(N), C.c_int(K), C.c_int(n), src, ln, dst,
                                     ptr(ws), C.c_uint64(nbytes), stream_ptr())
    check(st, "upnerf_wgrad_bf16_det")


def gemm_f32(A, sa, B, sb, C_out, sc, M, N, K, ep: Epilogue | None = None, accumulate=False,
             split_k=1):
    """Strided fp32 SIMT GEMM (see upnerf_gemm_f32). sa=(sam,sak), sb=(sbn,sbk), sc=(scm,scn)."""
    st = lib().upnerf_gemm_f32(ptr(A), _i64(sa[0]), _i64(sa[1]), ptr(B), _i64(sb[0]), _i64(sb[1]),
                               ptr(C_out), _i64(sc[0]), _i64(sc[1]), _i64(M), _i64(N), _i64(K),
                               C.byref(ep) if ep is not None else None, C.c_int(int(accumulate)),
                               C.c_int(split_k), stream_ptr())
    check(st, "upnerf_gemm_f32")


def gemm_tf32(A, sa, B, sb, C_out, sc, M, N, K, ep: Epilogue | None = None, accumulate=False, split_k=1):
    """The same strided product on tcgen05 with tf32 operands (see upnerf_gemm_tf32)."""
    st = lib().upnerf_gemm_tf32(ptr(A), _i64(sa[0]), _i64(sa[1]), ptr(B), _i64(sb[0]), _i64(sb[1]),
                                ptr(C_out), _i64(sc[0]), _i64(sc[1]), _i64(M), _i64(N), _i64(K),
                                C.byref(ep) if ep is not None else None, C.c_int(int(accumulate)),
                                C.c_int(split_k), stream_ptr())
    check(st, "upnerf_gemm_tf32")


# ---------------------------------------------------------------------------------------
# Structs of the render-level ABI (mirror include/upnerf_b200.h field for field)
# ---------------------------------------------------------------------------------------
F32, BF16 = 0, 1


class NetConfig(C.Structure):
    _fields_ = [("D", C.c_int), ("W", C.c_int), ("xyz_L", C.c_int), ("dir_L", C.c_int),
                ("encode_feat", C.c_int), ("feat_dim", C.c_int), ("appearance_dim", C.c_int),
                ("candidate_dim", C.c_int), ("encode_appearance", C.c_int), ("encode_candidate", C.c_int),
                ("use_c2f", C.c_int), ("c2f_start", C.c_float), ("c2f_end", C.c_float)]


_PASS_PTRS = ["params", "emb_a", "emb_c",
              "c_weights", "s_weights", "c_depth", "s_depth", "t_weight", "feat", "s_rgb",
              "g_c_weights", "g_s_weights", "g_c_depth", "g_s_depth", "g_t_weight", "g_feat", "g_s_rgb",
              "d_params", "d_emb_a", "d_emb_c"]


class PassIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _PASS_PTRS]


class RenderArgs(C.Structure):
    _fields_ = [("cfg", NetConfig), ("dtype", C.c_int), ("n_rays", C.c_int64), ("n_samples", C.c_int),
                ("n_importance", C.c_int), ("n_importance_static", C.c_int), ("n_images", C.c_int),
                ("sched_mult", C.c_float), ("use_disp", C.c_int), ("perturb", C.c_float),
                ("rays", C.c_void_p), ("img_idx", C.c_void_p), ("perturb_rand", C.c_void_p),
                ("u0", C.c_void_p), ("u1", C.c_void_p), ("coarse", PassIO), ("fine", PassIO),
                ("z_coarse", C.c_void_p), ("z_fine", C.c_void_p), ("d_rays", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("no_grad", C.c_int),
                ("reuse_packed", C.c_int)]


_TAIL_IN = ("img_idx", "inv_depths", "depth_scale", "rgbs", "feats", "s_depth_c", "s_depth_f", "t_weight_c",
            "t_weight_f", "feat_c", "feat_f", "s_rgb_c", "s_rgb_f", "t_beta", "t_alpha")
_TAIL_OUT = ("losses", "g_s_depth_c", "g_s_depth_f", "g_feat_c", "g_feat_f", "g_s_rgb_c", "g_s_rgb_f", "g_t_beta",
             "g_t_alpha", "d_depth_scale")
TAIL_LOSS_SLOTS = 16
TAIL_TERMS = ("l_depth_c", "l_feat_c", "l_rgb_c", "l_depth_f", "l_feat_f", "l_rgb_f", "l_beta", "l_alpha")


class TailArgs(C.Structure):
    """Mirror of `upnerf_tail_args`."""

    _fields_ = ([("n_rays", C.c_int64), ("feat_dim", C.c_int), ("has_fine", C.c_int), ("sched_mult", C.c_float),
                 ("depth_mult", C.c_float), ("alpha_reg", C.c_float), ("near_", C.c_float), ("far_", C.c_float)]
                + [(n, C.c_void_p) for n in _TAIL_IN + _TAIL_OUT]
                + [("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("sched_mult_dev", C.c_void_p)])


TNET_PARAMS = 19


class TnetArgs(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_images", C.c_int), ("feat_dim", C.c_int), ("hidden", C.c_int),
                ("transient_dim", C.c_int), ("beta_min", C.c_float), ("feats", C.c_void_p), ("img_idx", C.c_void_p),
                ("params", C.c_void_p * TNET_PARAMS), ("d_params", C.c_void_p * TNET_PARAMS),
                ("alpha", C.c_void_p), ("beta", C.c_void_p), ("rgb", C.c_void_p),
                ("g_alpha", C.c_void_p), ("g_beta", C.c_void_p), ("g_rgb", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


def tnet_workspace_bytes(a: TnetArgs) -> int:
    f = lib().upnerf_tnet_workspace_bytes
    f.restype = C.c_uint64
    n = int(f(C.byref(a)))
    if n == 0:
        raise UpnerfError("tnet: " + lib().upnerf_last_error().decode("utf-8", "replace"))
    return n


def tnet_fwd(a: TnetArgs) -> None:
    check(lib().upnerf_tnet_fwd(C.byref(a), stream_ptr()), "upnerf_tnet_fwd")


def tnet_bwd(a: TnetArgs) -> None:
    check(lib().upnerf_tnet_bwd(C.byref(a), stream_ptr()), "upnerf_tnet_bwd")


def tail_workspace_bytes() -> int:
    f = lib().upnerf_tail_workspace_bytes
    f.restype = C.c_uint64
    return int(f())


def tail_loss(a: TailArgs):
    check(lib().upnerf_tail_loss(C.byref(a), stream_ptr()), "upnerf_tail_loss")


ADAM_MAX_SEGMENTS = 64


class AdamArgs(C.Structure):
    """Mirror of `upnerf_adam_args`."""

    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("n_segments", C.c_int),
                ("seg_end", C.c_int64 * ADAM_MAX_SEGMENTS), ("seg_step_size", C.c_float * ADAM_MAX_SEGMENTS),
                ("seg_bc2_sqrt", C.c_float * ADAM_MAX_SEGMENTS), ("seg_live", C.c_int * ADAM_MAX_SEGMENTS),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("decay_mul", C.c_double),
                ("dev_scalars", C.c_void_p)]


def adam_step(a: AdamArgs):
    check(lib().upnerf_adam_step(C.byref(a), stream_ptr()), "upnerf_adam_step")


class CompositeArgs(C.Structure):
    _fields_ = ([("R", C.c_int64), ("S", C.c_int), ("cand", C.c_int), ("stat_rgb", C.c_int),
                 ("feat_mode", C.c_int), ("dtype", C.c_int)]
                + [(n, C.c_void_p) for n in ("z", "s_sigma", "c_sigma", "rgb", "hf")]
                + [("ld_hf", C.c_int64), ("g2", C.c_void_p), ("ld_g2", C.c_int64)]
                + [(n, C.c_void_p) for n in ("c_weights", "s_weights", "c_depth", "t_weight", "s_depth", "s_rgb",
                                             "hf_ray", "g2_ray", "ws_sum", "wc_sum",
                                             "g_c_weights", "g_s_weights", "g_c_depth", "g_t_weight", "g_s_depth",
                                             "g_s_rgb", "g_hf_ray", "g_g2_ray", "g_ws_sum", "g_wc_sum", "w_csigma",
                                             "d_ssig_pre", "d_csig_pre", "d_rgb", "d_hf")]
                + [("ld_dhf", C.c_int64), ("d_g2pre", C.c_void_p), ("ld_dg2", C.c_int64)])


TRUNK_LAYERS, TRUNK_WCAT_COLS = 9, 2176


class TrunkArgs(C.Structure):
    """Mirror of `upnerf_trunk_args`."""

    _fields_ = [("pe", C.c_void_p), ("ld_pe", C.c_int64), ("wcat", C.c_void_p), ("ld_w", C.c_int64),
                ("bias", C.c_void_p * TRUNK_LAYERS), ("sigma_w", C.c_void_p), ("sigma_b", C.c_void_p),
                ("out", C.c_void_p * TRUNK_LAYERS), ("ld_out", C.c_int64 * TRUNK_LAYERS),
                ("s_sigma", C.c_void_p), ("M", C.c_int64), ("relu_mask", C.c_void_p)]


TRUNK_BWD_LAYERS, TRUNK_WCATT_COLS = 8, 2048


class TrunkBwdArgs(C.Structure):
    """Mirror of `upnerf_trunk_bwd_args`."""

    _fields_ = [("d_hf", C.c_void_p), ("ld_dhf", C.c_int64), ("d_ssig", C.c_void_p), ("sigma_w", C.c_void_p),
                ("wcat_t", C.c_void_p), ("ld_w", C.c_int64), ("relu_mask", C.c_void_p),
                ("d_out", C.c_void_p * TRUNK_BWD_LAYERS), ("ld_dout", C.c_int64 * TRUNK_BWD_LAYERS),
                ("M", C.c_int64)]


def trunk_mask_words(M: int) -> int:
    f = lib().upnerf_trunk_mask_words
    f.restype = C.c_int64
    return int(f(_i64(M)))


def mlp_trunk_bwd(d_hf, d_ssig, sigma_w, wcat_t, relu_mask, d_outs, M):
    """Fused backward data-gradient chain (see upnerf_mlp_trunk_bwd_bf16): d_outs = [dY8, ..., dY1]."""
    a = TrunkBwdArgs()
    a.d_hf, a.ld_dhf = ptr(d_hf).value, d_hf.stride(0)
    a.d_ssig = None if d_ssig is None else ptr(d_ssig).value
    a.sigma_w = ptr(sigma_w).value
    a.wcat_t, a.ld_w = ptr(wcat_t).value, wcat_t.stride(0)
    a.relu_mask = ptr(relu_mask).value
    for j in range(TRUNK_BWD_LAYERS):
        a.d_out[j] = ptr(d_outs[j]).value
        a.ld_dout[j] = d_outs[j].stride(0)
    a.M = int(M)
    check(lib().upnerf_mlp_trunk_bwd_bf16(C.byref(a), stream_ptr()), "upnerf_mlp_trunk_bwd_bf16")


def mlp_trunk_fwd(pe, wcat, biases, sigma_w, sigma_b, outs, s_sigma, M, relu_mask=None):
    """Fused trunk forward (see upnerf_mlp_trunk_fwd_bf16): pe [M,64], wcat [256,2176] bf16,
    biases: 9 fp32 [256], outs: 9 bf16 [M,256] (any row stride), s_sigma [M] fp32."""
    a = TrunkArgs()
    a.pe, a.ld_pe = pe.data_ptr(), pe.stride(0)
    a.wcat, a.ld_w = wcat.data_ptr(), wcat.stride(0)
    for i in range(TRUNK_LAYERS):
        ptr(biases[i]), ptr(outs[i])          # CUDA-only check
        a.bias[i] = biases[i].data_ptr()
        a.out[i] = outs[i].data_ptr()
        a.ld_out[i] = outs[i].stride(0)
    a.sigma_w, a.sigma_b, a.s_sigma = sigma_w.data_ptr(), sigma_b.data_ptr(), s_sigma.data_ptr()
    a.M = int(M)
    a.relu_mask = None if relu_mask is None else ptr(relu_mask).value
    check(lib().upnerf_mlp_trunk_fwd_bf16(C.byref(a), stream_ptr()), "upnerf_mlp_trunk_fwd_bf16")


def _vp(t):
    return None if t is None else t.data_ptr()


def nerf_param_count(cfg: NetConfig) -> int:
    f = lib().upnerf_nerf_param_count
    f.restype = C.c_int64
    n = f(C.byref(cfg))
    if n < 0:
        check(2, "upnerf_nerf_param_count")
    return int(n)


def render_workspace_bytes(args: RenderArgs) -> int:
    f = lib().upnerf_render_workspace_bytes
    f.restype = C.c_uint64
    n = int(f(C.byref(args)))
    if n == 0:
        check(1, "upnerf_render_workspace_bytes")
    return n


def render_fwd(args: RenderArgs) -> None:
    check(lib().upnerf_render_fwd(C.byref(args), stream_ptr()), "upnerf_render_fwd")


def render_bwd_passes(args: RenderArgs, passes: int) -> None:
    """passes: bit 0 = fine network, bit 1 = coarse network."""
    check(lib().upnerf_render_bwd_passes(C.byref(args), C.c_int(passes), stream_ptr()), "upnerf_render_bwd_passes")


def render_bwd(args: RenderArgs) -> None:
    check(lib().upnerf_render_bwd(C.byref(args), stream_ptr()), "upnerf_render_bwd")


def composite_fwd(args: CompositeArgs) -> None:
    check(lib().upnerf_composite_fwd(C.byref(args), stream_ptr()), "upnerf_composite_fwd")


def composite_bwd(args: CompositeArgs) -> None:
    check(lib().upnerf_composite_bwd(C.byref(args), stream_ptr()), "upnerf_composite_bwd")


def pose_rays_fwd(table, img_idx, c2w, directions, near_far, rays, pose_out=None):
    single = 1 if c2w.dim() == 2 else 0
    st = lib().upnerf_pose_rays_fwd(ptr(table), ptr(img_idx), ptr(c2w), C.c_int(single), ptr(directions),
                                    ptr(near_far), _i64(directions.shape[0]), ptr(rays), ptr(pose_out),
                                    stream_ptr())
    check(st, "upnerf_pose_rays_fwd")


def pose_rays_bwd(table, img_idx, c2w, directions, d_rays, d_table):
    single = 1 if c2w.dim() == 2 else 0
    st = lib().upnerf_pose_rays_bwd(ptr(table), ptr(img_idx), ptr(c2w), C.c_int(single), ptr(directions),
                                    _i64(directions.shape[0]), ptr(d_rays), ptr(d_table), stream_ptr())
    check(st, "upnerf_pose_rays_bwd")


def stratified_z(rays, perturb_rand, perturb, use_disp, S, z):
    st = lib().upnerf_stratified_z(ptr(rays), ptr(perturb_rand), C.c_float(perturb), C.c_int(int(use_disp)),
                                   _i64(rays.shape[0]), C.c_int(S), ptr(z), stream_ptr())
    check(st, "upnerf_stratified_z")


def sample_pdf(bins, weights, u, N, eps, samples, inds=None, cdf_out=None):
    R, nw = weights.shape
    st = lib().upnerf_sample_pdf(ptr(bins), _i64(bins.stride(0)), ptr(weights), _i64(weights.stride(0)), ptr(u),
                                 _i64(R), C.c_int(nw), C.c_int(N), C.c_float(eps), ptr(samples), ptr(inds),
                                 ptr(cdf_out), stream_ptr())
    check(st, "upnerf_sample_pdf")


def searchsorted_right(cdf, u, inds):
    st = lib().upnerf_searchsorted_right(ptr(cdf), C.c_int(cdf.shape[1]), ptr(u), C.c_int(u.shape[1]),
                                         _i64(cdf.shape[0]), ptr(inds), stream_ptr())
    check(st, "upnerf_searchsorted_right")


def resample_merge(z, w0, w1, ld_w, u0, u1, n0, n1, eps, z_fine):
    R, S = z.shape
    st = lib().upnerf_resample_merge(ptr(z), ptr(w0), ptr(w1), _i64(ld_w), ptr(u0), ptr(u1), C.c_int(n0),
                                     C.c_int(n1), _i64(R), C.c_int(S), C.c_float(eps), ptr(z_fine), stream_ptr())
    check(st, "upnerf_resample_merge")


def c2f_weights(progress_dev, start, end, use_c2f, L, out):
    st = lib().upnerf_c2f_weights(ptr(progress_dev), C.c_float(start), C.c_float(end), C.c_int(int(use_c2f)),
                                  C.c_int(L), ptr(out), stream_ptr())
    check(st, "upnerf_c2f_weights")


def posenc_fwd(x, ld_x, M, L, band_w, out, ld_out, width, dtype):
    st = lib().upnerf_posenc_fwd(ptr(x), _i64(ld_x), _i64(M), C.c_int(L), ptr(band_w), ptr(out), _i64(ld_out),
                                 C.c_int(width), C.c_int(dtype), stream_ptr())
    check(st, "upnerf_posenc_fwd")


def points_posenc_fwd(rays, z, L, band_w, out, ld_out, width, dtype):
    R, S = z.shape
    st = lib().upnerf_points_posenc_fwd(ptr(rays), ptr(z), _i64(R), C.c_int(S), C.c_int(L), ptr(band_w), ptr(out),
                                        _i64(ld_out), C.c_int(width), C.c_int(dtype), stream_ptr())
    check(st, "upnerf_points_posenc_fwd")


def points_posenc_bwd(d_pe, ld_pe, rays, z, L, band_w, d_rays, dtype):
    R, S = z.shape
    st = lib().upnerf_points_posenc_bwd(ptr(d_pe), _i64(ld_pe), ptr(rays), ptr(z), _i64(R), C.c_int(S), C.c_int(L),
                                        ptr(band_w), ptr(d_rays), C.c_int(dtype), stream_ptr())
    check(st, "upnerf_points_posenc_bwd")


def se3_exp_fwd(wu, out):
    check(lib().upnerf_se3_exp_fwd(ptr(wu), _i64(wu.shape[0]), ptr(out), stream_ptr()), "upnerf_se3_exp_fwd")


def se3_exp_bwd(wu, d_pose, d_wu):
    check(lib().upnerf_se3_exp_bwd(ptr(wu), ptr(d_pose), _i64(wu.shape[0]), ptr(d_wu), stream_ptr()),
          "upnerf_se3_exp_bwd")


def pose_compose_fwd(a, a_single, b, b_single, n, out):
    check(lib().upnerf_pose_compose_fwd(ptr(a), C.c_int(int(a_single)), ptr(b), C.c_int(int(b_single)), _i64(n),
                                        ptr(out), stream_ptr()), "upnerf_pose_compose_fwd")


def pose_compose_bwd(a, a_single, b, b_single, d_out, n, d_a, d_b):
    check(lib().upnerf_pose_compose_bwd(ptr(a), C.c_int(int(a_single)), ptr(b), C.c_int(int(b_single)), ptr(d_out),
                                        _i64(n), ptr(d_a), ptr(d_b), stream_ptr()), "upnerf_pose_compose_bwd")


def get_rays_bwd(c2w, directions, d_rays, d_c2w):
    single = 1 if c2w.dim() == 2 else 0
    check(lib().upnerf_get_rays_bwd(ptr(c2w), C.c_int(single), ptr(directions), _i64(directions.shape[0]),
                                    ptr(d_rays), ptr(d_c2w), stream_ptr()), "upnerf_get_rays_bwd")


_RAY_BATCH_PTRS = ("idx", "ray_infos", "directions", "rgbs", "pxl_coords", "inv_depths", "feat_maps", "poses",
                   "out_ray_infos", "out_directions", "out_img_idx", "out_c2w", "out_rgbs", "out_feats",
                   "out_inv_depths", "status")


class RayBatchArgs(C.Structure):
    """Mirror of `upnerf_ray_batch_args`."""

    _fields_ = ([("n_rays", C.c_int64), ("n_total", C.c_int64), ("n_images", C.c_int), ("feat_h", C.c_int),
                 ("feat_w", C.c_int), ("feat_dim", C.c_int)]
                + [(n, C.c_void_p) for n in _RAY_BATCH_PTRS])


def ray_batch_gather(a: RayBatchArgs):
    check(lib().upnerf_ray_batch_gather(C.byref(a), stream_ptr()), "upnerf_ray_batch_gather")
