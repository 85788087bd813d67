"""HBM rate of the resident ray batcher at the reference's feature-map shape.

    python tools/bench_batcher.py [--images 256] [--rays 4096 32768 262144]

Tables: --images synthetic images of 111 x 111 x 384 fp32 features (18.9 MB each; the full
763-image scene is 14.4 GB and also fits) and 64 x 48 rays per image.  Reports, per batch size, the
device time of one gather (CUDA events, median of 20 after 5 warm-ups, a fresh random index set each
time so the feature rows come from HBM, not L2) and algorithmic bytes / time against the measured HBM
peak; next to it the oracle (the reference's per-ray arithmetic, vectorised) on the host cores."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--rays", type=int, nargs="+", default=[4096, 32768, 262144])
    args = ap.parse_args()
    from oracle import ray_batch as RB
    from upnerf_b200.datasets import RayBatcher

    dev = torch.device("cuda:0")
    n_img, fh, C, ih, iw = args.images, 111, 384, 48, 64
    g = torch.Generator(device=dev).manual_seed(0)
    N = n_img * ih * iw
    fm = torch.randn(n_img, fh, fh, C, device=dev, generator=g)
    fm /= fm.norm(dim=-1, keepdim=True)
    infos = torch.stack([torch.full((N,), 0.1, device=dev), torch.full((N,), 5.0, device=dev),
                         torch.arange(N, device=dev).div(ih * iw, rounding_mode="floor").float()], 1)
    b = RayBatcher(infos, torch.randn(N, 3, device=dev, generator=g), torch.rand(N, 3, device=dev, generator=g),
                   torch.eye(3, 4, device=dev).expand(n_img, 3, 4), all_pxl_coords=torch.rand(N, 2, device=dev, generator=g),
                   feat_maps=fm, all_inv_depths=torch.rand(N, device=dev, generator=g), device=dev)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6550.4))
    per_ray = 5 * C * 4 + 8 + 8 + 2 * (12 + 12 + 8 + 48) + 12 + 8 + 8     # = include/upnerf_b200.h (f1)
    for R in args.rays:
        idxs = [torch.randint(0, N, (R,), device=dev, generator=g) for _ in range(25)]
        for i in range(5):
            b.gather(idxs[i])
        ts = []
        for i in range(5, 25):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.gather(idxs[i])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbs = R * per_ray / (ms * 1e-3) / 1e9
        line = {"kernel": "ray_batch_gather_kernel", "rays": R, "images": n_img, "ms": round(ms, 4),
                "rays_per_s": round(R / (ms * 1e-3)), "algorithmic_bytes_per_ray": per_ray,
                "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}
        if R <= 32768:
            tabs = {"all_ray_infos": b.all_ray_infos.cpu(), "all_directions": b.all_directions.cpu(),
                    "all_rgbs": b.all_rgbs.cpu(), "all_pxl_coords": b.all_pxl_coords.cpu(),
                    "all_inv_depths": b.all_inv_depths.cpu(), "feat_maps": b.feat_maps.cpu(), "poses": b.poses.cpu()}
            idx = idxs[0].cpu()
            torch.set_num_threads(os.cpu_count())
            RB.getitem_batch(tabs, idx)
            t0 = time.perf_counter()
            for _ in range(3):
                RB.getitem_batch(tabs, idx)
            cpu_s = (time.perf_counter() - t0) / 3
            line["cpu_oracle_rays_per_s"] = round(R / cpu_s)
            line["cpu_cores"] = os.cpu_count()
            del tabs
        print(json.dumps(line), flush=True)
    b.check()


if __name__ == "__main__":
    main()
