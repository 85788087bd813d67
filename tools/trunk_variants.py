"""Timing A/B of the fused trunk kernels (CTA pairs vs single CTAs, masks on / off) in one process (CUDA events, M = 786k).
usage: python tools/trunk_variants.py [NAME=ENV1:VAL,ENV2:VAL ...]   (default: a fixed list of variants)"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upnerf_b200 import _lib as L
from tools.bench_gemm import timeit

VARIANTS = [
    ("CTA pairs (default)", {}),
    ("single CTAs", {"UPNERF_TRUNK_CLUSTER": "1"}),
    ("CTA pairs, no masks", {"TV_NOMASK": "1"}),
]
KEYS = ("TV_NOMASK", "UPNERF_TRUNK_CLUSTER")


def main():
    dev = torch.device("cuda:0")
    M = 4096 * 192
    variants = VARIANTS
    if len(sys.argv) > 1:
        variants = []
        for a in sys.argv[1:]:
            name, _, env = a.partition("=")
            variants.append((name, dict(kv.split(":") for kv in env.split(",") if kv)))
    pe = torch.randn(M, 64, device=dev).bfloat16()
    ks = [64, 256, 256, 256, 320, 256, 256, 256, 256]
    wcat = torch.cat([(torch.randn(256, k, device=dev) * (1.4 / k ** 0.5)).bfloat16() for k in ks], 1).contiguous()
    bs = [torch.randn(256, device=dev) * 0.05 for _ in ks]
    sw, sb = torch.randn(256, device=dev) / 16, torch.zeros(1, device=dev)
    outs = [torch.empty(M, 256, device=dev, dtype=torch.bfloat16) for _ in ks]
    sig = torch.empty(M, device=dev)
    mask = torch.zeros(L.trunk_mask_words(M), dtype=torch.int32, device=dev)
    wcat_t = (torch.randn(256, 2048, device=dev) / 16).bfloat16()
    d_hf = torch.randn(M, 256, device=dev).bfloat16()
    d_ssig = torch.randn(M, device=dev)
    d_outs = [torch.empty(M, 256, device=dev, dtype=torch.bfloat16) for _ in range(8)]
    # the variants are interleaved and repeated (the box's clocks drift under sustained load: a single pass
    # over the list favours whatever runs first); min and median over the rounds are reported
    res = {name: ([], []) for name, _ in variants}
    for _ in range(5):
        for name, env in variants:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            res[name][0].append(timeit(lambda: L.mlp_trunk_fwd(pe, wcat, bs, sw, sb, outs, sig, M, relu_mask=None if os.environ.get('TV_NOMASK') else mask), iters=8, warm=1))
            res[name][1].append(timeit(lambda: L.mlp_trunk_bwd(d_hf, d_ssig, sw, wcat_t, mask, d_outs, M), iters=8, warm=1))
    for name, (f, b) in res.items():
        f, b = sorted(f), sorted(b)
        print(f"{name:40s} fwd(+mask) min {f[0]:.3f} med {f[len(f) // 2]:.3f} ms   bwd min {b[0]:.3f} med {b[len(b) // 2]:.3f} ms", flush=True)
    for k in KEYS:
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
