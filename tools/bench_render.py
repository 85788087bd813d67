"""BASELINE config 5: chunked full-image inference render, 1920x1080 = 2,073,600 rays, coarse+fine
(64+64), phase 2 (sched_mult = 1, the tto / validation path of models/nerf_system_optmize.py:84-111),
perturb = 0, under torch.no_grad.  Reports rays/s and seconds per image for the reference's chunk
size (val.chunk_size = 4096) and for larger chunks.

    python tools/bench_render.py [--chunks 4096,65536] [--precision bf16]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

MACS_PHASE2 = 714_240          # per-sample MACs of the reference MLP in phase 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", default="4096,32768,131072")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    from upnerf_b200.models.nerf_system import NeRFSystem
    from upnerf_b200.utils.ray import get_ray_directions, get_rays

    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    hp = {"nerf.N_samples": 64, "nerf.N_importance": 64, "kernel.precision": args.precision}
    system = NeRFSystem(hp, N_images_train=763, device=dev)
    system.set_progress(0.75)
    H, Wd = args.height, args.width
    K = torch.tensor([[0.8 * Wd, 0, Wd / 2], [0, 0.8 * Wd, H / 2], [0, 0, 1]])
    dirs = get_ray_directions(H, Wd, K).to(dev)
    c2w = torch.eye(3, 4, device=dev)
    o, d = get_rays(dirs, c2w)
    R = H * Wd
    rays = torch.cat([o.reshape(-1, 3), d.reshape(-1, 3), torch.full((R, 1), 0.1, device=dev),
                      torch.full((R, 1), 5.0, device=dev)], 1)
    idx = torch.zeros(R, dtype=torch.long, device=dev)
    feats = torch.zeros(R, 384, device=dev)
    peaks = ROOT / "MEASURED_PEAKS.json"
    tf_peak = json.loads(peaks.read_text())["bf16_tflops_sustained"] if peaks.exists() else 1400.0
    for chunk in [int(c) for c in args.chunks.split(",")]:
        system.hparams["val.chunk_size"] = chunk
        with torch.no_grad():
            # warm-up: ~0.25 M rays through the same chunk size (a fresh process's first launches, the
            # caching allocator and the clocks settle; two chunks were not enough for 4096-ray chunks)
            n_w = max(chunk * 2, min(R, 262144))
            res = system(rays[:n_w], feats[:n_w], idx[:n_w], 1.0, train=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = system(rays, feats, idx, 1.0, train=False)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        assert res["rgb_fine"].shape == (R, 3) and torch.isfinite(res["rgb_fine"]).all()
        flops = R * 192 * MACS_PHASE2 * 2
        print(json.dumps({"workload": f"config 5: {Wd}x{H} inference render, 64+64 samples, phase 2, chunk {chunk}",
                          "precision": args.precision, "rays": R, "ms_per_image": round(ms, 2),
                          "rays_per_s": round(R / ms * 1e3), "algorithmic_tflops": round(flops / ms / 1e9, 1),
                          "frac_of_sustained_bf16_peak": round(flops / ms / 1e9 / tf_peak, 3),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))


if __name__ == "__main__":
    main()
