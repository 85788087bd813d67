"""HBM-roofline measurement of the streaming kernels (compositing, sample_pdf + merge, positional
encoding) at sizes far beyond the 126 MB L2 (at 4096-8192 rays they are
L2-resident and launch-bound, so the HBM fraction is only meaningful at >= 2^17 rays).

Prints one JSON line per kernel: algorithmic bytes (every operand touched once), CUDA-event time,
GB/s and the fraction of the measured HBM peak (MEASURED_PEAKS.json).  `--only NAME` restricts the
run to one kernel (for `ncu -k regex:NAME`).
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from upnerf_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def report(name, nbytes, ms, peak, note):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "algorithmic_bytes": nbytes, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / peak, 3), "note": note}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(pk.read_text())["hbm_gbs"] if pk.exists() else 6650.0
    want = lambda n: not args.only or args.only in n
    g = torch.Generator(device="cuda").manual_seed(0)

    # ---------------------------------------------------------------- sample_pdf + sort-merge
    if want("resample_merge"):
        R, S, NI = 1 << 20, 64, 64
        z = torch.sort(0.1 + 4.9 * torch.rand(R, S, device=dev, generator=g), -1)[0]
        w = torch.rand(R, S, device=dev, generator=g) ** 4
        u = torch.rand(R, NI, device=dev, generator=g)
        zf = torch.empty(R, S + NI, device=dev)
        ms = timeit(lambda: L.resample_merge(z, w[:, 1:], None, S, u, None, NI, 0, 1e-5, zf), args.iters)
        nbytes = R * 4 * (S + (S - 2) + NI + (S + NI))       # z, weights[1:-1], u in; z_fine out (1,524 B/ray)
        report("resample_merge_kernel", nbytes, ms, peak, f"R=2^20 rays, {S}+{NI} samples, single draw")

    # ---------------------------------------------------------------- compositing (phase 1: all outputs)
    R, S = 1 << 17, 128
    M = R * S
    if want("composite"):
        a = L.CompositeArgs()
        a.R, a.S, a.cand, a.stat_rgb, a.feat_mode, a.dtype = R, S, 1, 1, 2, L.BF16
        t = dict(z=torch.sort(0.1 + 4.9 * torch.rand(R, S, device=dev, generator=g), -1)[0],
                 s_sigma=torch.rand(M, device=dev, generator=g) * 2, c_sigma=torch.rand(M, device=dev, generator=g),
                 rgb=torch.rand(M, 3, device=dev, generator=g),
                 hf=torch.randn(M, 256, device=dev, generator=g).bfloat16(),
                 g2=torch.randn(M, 128, device=dev, generator=g).clamp_min(0).bfloat16())
        for k, v in t.items():
            setattr(a, k, v.data_ptr())
        a.ld_hf, a.ld_g2 = 256, 128
        outs = dict(c_weights=(R, S), s_weights=(R, S), c_depth=(R,), t_weight=(R,), s_depth=(R,), s_rgb=(R, 3),
                    hf_ray=(R, 256), g2_ray=(R, 128), ws_sum=(R,), wc_sum=(R,))
        to = {k: torch.empty(s, device=dev) for k, s in outs.items()}
        for k, v in to.items():
            setattr(a, k, v.data_ptr())
        if want("composite_fwd"):
            ms = timeit(lambda: L.composite_fwd(a), args.iters)
            per_sample = 4 * 3 + 12 + 2 * (256 + 128) + 8      # z, 2 sigmas, rgb, HF + G2 (bf16) in; 2 weights out
            nbytes = M * per_sample + R * 4 * (6 + 256 + 128 + 2)
            report("composite_fwd_kernel", nbytes, ms, peak, f"R=2^17 rays x {S} samples, phase 1, hidden-vector compositing")
        if want("composite_bwd"):
            gi = {k: torch.randn(s, device=dev, generator=g) for k, s in
                  dict(g_c_weights=(R, S), g_s_weights=(R, S), g_c_depth=(R,), g_t_weight=(R,), g_s_depth=(R,),
                       g_s_rgb=(R, 3), g_hf_ray=(R, 256), g_g2_ray=(R, 128), g_ws_sum=(R,), g_wc_sum=(R,)).items()}
            for k, v in gi.items():
                setattr(a, k, v.data_ptr())
            wcs = torch.randn(128, device=dev, generator=g) * 0.1
            a.w_csigma = wcs.data_ptr()
            do = dict(d_ssig_pre=torch.empty(M, device=dev), d_csig_pre=torch.empty(M, device=dev),
                      d_rgb=torch.empty(M, 3, device=dev), d_hf=torch.empty(M, 256, device=dev, dtype=torch.bfloat16),
                      d_g2pre=torch.empty(M, 128, device=dev, dtype=torch.bfloat16))
            for k, v in do.items():
                setattr(a, k, v.data_ptr())
            a.ld_dhf, a.ld_dg2 = 256, 128
            ms = timeit(lambda: L.composite_bwd(a), args.iters)
            per_sample = (4 * 3 + 12 + 2 * (256 + 128)) + 8 + (4 * 2 + 12 + 2 * (256 + 128))
            nbytes = M * per_sample + R * 4 * (6 + 256 + 128 + 2)
            report("composite_bwd_kernel", nbytes, ms, peak, f"R=2^17 rays x {S} samples, phase 1 (reads fwd inputs + 2 weight grads, writes 5 gradients)")

    # ---------------------------------------------------------------- positional encoding of x = o + d z
    if want("posenc"):
        rays = torch.randn(R, 8, device=dev, generator=g)
        z = torch.rand(R, S, device=dev, generator=g)
        band = torch.ones(16, device=dev)
        out = torch.empty(M, 64, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: L.points_posenc_fwd(rays, z, 10, band, out, 64, 64, L.BF16), args.iters)
        report("points_posenc_fwd_kernel", M * (4 + 128) + R * 32, ms, peak, f"R=2^17 x {S}: read z, write 64 bf16 per sample")
        d_pe = torch.randn(M, 64, device=dev, generator=g).bfloat16()
        d_rays = torch.zeros(R, 8, device=dev)
        ms = timeit(lambda: L.points_posenc_bwd(d_pe, 64, rays, z, 10, band, d_rays, L.BF16), args.iters)
        report("points_posenc_bwd_kernel", M * (4 + 128) + R * 64, ms, peak, f"R=2^17 x {S}: read z + dPE, reduce to d_rays")


if __name__ == "__main__":
    main()
