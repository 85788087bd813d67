"""Micro-benchmark of the dense-layer primitives (CUDA events, L2-exceeding inputs)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upnerf_b200 import _lib as L


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


def main():
    dev = torch.device("cuda:0")
    M = 4096 * 192
    for a in sys.argv[1:]:
        if a.startswith("M="):
            M = int(a[2:])
    only_trunk = "trunk" in sys.argv[1:]
    if "epi" in sys.argv[1:]:       # which epilogue feature costs what (the head layers of the render forward)
        Mm, S = 4096 * 128, 128
        for (N, K) in [(256, 256), (128, 128)]:
            A = torch.randn(Mm, 256, device=dev).bfloat16()[:, :K]
            B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
            Cc = torch.empty(Mm, 256, device=dev, dtype=torch.bfloat16)[:, :N]
            bias = torch.randn(N, device=dev)
            rb = torch.randn(Mm // S, N, device=dev)
            hw = torch.randn(3, N, device=dev)
            hb = torch.zeros(3, device=dev)
            ho = torch.empty(Mm, 3, device=dev)
            variants = {
                "bias+relu": dict(bias=bias, act=1),
                "ray_bias+relu": dict(ray_bias=rb, rows_per_ray=S, act=1),
                "bias+relu+3 heads": dict(bias=bias, act=1, head_w=hw, head_b=hb, head_act=2, head_out=ho),
                "ray_bias+relu+3 heads": dict(ray_bias=rb, rows_per_ray=S, act=1, head_w=hw, head_b=hb, head_act=2, head_out=ho),
                "ray_bias+relu+1 head": dict(ray_bias=rb, rows_per_ray=S, act=1, head_w=hw[:1].contiguous(), head_b=hb, head_act=1, head_out=ho),
            }
            for name, kw in variants.items():
                ep = L.make_epilogue(**kw)
                if "3 heads" in name and N == 256:
                    ep.head_col_begin = 128
                ms = timeit(lambda: L.gemm_bf16(A, B, Cc, Mm, N, K, ep=ep), iters=20)
                by = (Mm * K + Mm * N) * 2
                print(f"gemm_bf16 M={Mm} N={N} K={K} lda={A.stride(0)} {name:24s}: {ms * 1e3:7.1f} us  {by / ms / 1e6:6.0f} GB/s")
        return
    if "tf32" in sys.argv[1:]:      # the small per-ray / parameter-space products (render.cu), one launch each
        R = 4096
        for name, (Mm, Nn, Kk), sa, sb, split, acc in [
                ("feat = HFr Wsf^T", (R, 384, 256), (256, 1), (256, 1), 1, False),
                ("gHFr = gf Wsf", (R, 256, 384), (384, 1), (1, 256), 1, False),
                ("dWsf += gf^T HFr", (384, 256, R), (1, 384), (1, 256), 32, False),
                ("Bq = P W^T", (R, 128, 75), (75, 1), (459, 1), 1, False),
                ("Wq32 = Wr0 Wsf", (128, 256, 384), (459, 1), (1, 256), 1, False),
                ("gWr0 += dWq Wsf^T", (128, 384, 256), (256, 1), (256, 1), 8, True)]:
            A = torch.randn(max(Mm, Kk) * 512, device=dev)
            B = torch.randn(max(Nn, Kk) * 512, device=dev)
            Cc = torch.zeros(Mm, Nn, device=dev)
            for impl, fn in (("tf32", L.gemm_tf32), ("simt", L.gemm_f32)):
                ms = timeit(lambda: fn(A, sa, B, sb, Cc, (Nn, 1), Mm, Nn, Kk, accumulate=acc, split_k=split), iters=50)
                print(f"{impl} {name:22s} M={Mm} N={Nn} K={Kk} split={split}: {ms * 1e3:7.1f} us  "
                      f"{2.0 * Mm * Nn * Kk / ms / 1e9:6.2f} TFLOP/s")
        return
    for (N, K, aux) in [] if only_trunk else [(256, 256, 0), (256, 256, 2), (256, 64, 0), (256, 320, 0), (128, 256, 0), (64, 256, 0)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        Cc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev)
        auxt = torch.randn(M, N, device=dev).bfloat16() if aux else None
        ep = L.make_epilogue(bias=bias, act=1, aux=auxt, ldaux=N, aux_mode=aux)
        ms = timeit(lambda: L.gemm_bf16(A, B, Cc, M, N, K, ep=ep))
        fl = 2.0 * M * N * K
        by = (M * K + M * N * (2 if aux else 1)) * 2
        print(f"gemm_bf16 M={M} N={N} K={K} aux={aux}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: torch.matmul(A, B.t()))
        print(f"   torch.matmul bf16: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    for (N, K) in [] if only_trunk else [(256, 256), (256, 320), (128, 256), (256, 64)]:
        dY = torch.randn(M, N, device=dev).bfloat16()
        X = torch.randn(M, K, device=dev).bfloat16()
        dW = torch.zeros(N, K, device=dev)
        db = torch.zeros(N, device=dev)
        ms = timeit(lambda: L.wgrad_bf16(dY, X, dW, db, M, N, K, [(0, K, 0)]))
        fl = 2.0 * M * N * K
        by = (M * K + M * N) * 2
        print(f"wgrad_bf16 M={M} N={N} K={K}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: L.wgrad_bf16_det(dY, X, dW, db, M, N, K, [(0, K, 0)]))
        print(f"wgrad_bf16_det (partials + reduce launch) M={M} N={N} K={K}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s")
        ms = timeit(lambda: torch.matmul(dY.t(), X))
        print(f"   torch.matmul bf16: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    if "wgrad" in sys.argv[1:]:
        return
    # fused trunk forward
    pe = torch.randn(M, 64, device=dev).bfloat16()
    ks = [64, 256, 256, 256, 320, 256, 256, 256, 256]
    wcat = torch.cat([(torch.randn(256, k, device=dev) * (1.4 / k ** 0.5)).bfloat16() for k in ks], 1).contiguous()
    bs = [torch.zeros(256, device=dev) for _ in ks]
    sw, sb = torch.randn(256, device=dev) / 16, torch.zeros(1, device=dev)
    outs = [torch.empty(M, 256, device=dev, dtype=torch.bfloat16) for _ in ks]
    sig = torch.empty(M, device=dev)
    ms = timeit(lambda: L.mlp_trunk_fwd(pe, wcat, bs, sw, sb, outs, sig, M))
    fl = 2.0 * M * 256 * sum(ks)
    by = (M * 64 + 9 * M * 256) * 2
    print(f"mlp_trunk_fwd M={M}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s (write-only activations)")
    mask = torch.zeros(L.trunk_mask_words(M), dtype=torch.int32, device=dev)
    L.mlp_trunk_fwd(pe, wcat, bs, sw, sb, outs, sig, M, relu_mask=mask)
    ms = timeit(lambda: L.mlp_trunk_fwd(pe, wcat, bs, sw, sb, outs, sig, M, relu_mask=mask))
    print(f"mlp_trunk_fwd (+mask) M={M}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    wcat_t = (torch.randn(256, 2048, device=dev) / 16).bfloat16()
    d_hf = torch.randn(M, 256, device=dev).bfloat16()
    d_ssig = torch.randn(M, device=dev)
    d_outs = [torch.empty(M, 256, device=dev, dtype=torch.bfloat16) for _ in range(8)]
    ms = timeit(lambda: L.mlp_trunk_bwd(d_hf, d_ssig, sw, wcat_t, mask, d_outs, M))
    fl = 2.0 * M * 256 * 2048
    by = (M * 256 + 8 * M * 256) * 2 + mask.numel() * 4
    print(f"mlp_trunk_bwd M={M}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
    if only_trunk:
        return
    Mf = 256 * 192
    A = torch.randn(Mf, 256, device=dev)
    B = torch.randn(256, 256, device=dev)
    Cc = torch.empty(Mf, 256, device=dev)
    ms = timeit(lambda: L.gemm_f32(A, (256, 1), B, (256, 1), Cc, (256, 1), Mf, 256, 256))
    print(f"gemm_f32 M={Mf}: {ms:.3f} ms {2.0 * Mf * 65536 / ms / 1e9:.2f} TFLOP/s")


if __name__ == "__main__":
    main()
