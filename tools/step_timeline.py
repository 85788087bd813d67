"""Chronological device timeline of ONE train step (CUPTI via torch.profiler's chrome trace): per kernel its
stream, start offset, duration and the idle gap since the previous kernel on the same stream.

    python tools/step_timeline.py [--rays 4096] > gpurun_out/timeline.txt

Diagnostic only (numbers under a profiler are never bench values)."""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--min-us", type=float, default=0.0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    system = bench.make_system("bf16", dev)
    batches = [{k: v.to(dev) for k, v in bench.host_batch(args.rays, i, False).items()} for i in range(4)]
    for i in range(6):
        system.training_step(batches[i % 4], i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            system.training_step(batches[i % 4], i)
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "trace.json"
        prof.export_chrome_trace(str(path))
        tr = json.loads(path.read_text())
    evs = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    evs.sort(key=lambda e: e["ts"])
    # the middle step: from the second pose_rays_fwd kernel to the third
    marks = [i for i, e in enumerate(evs) if "pose_rays_fwd" in e["name"]]
    lo, hi = marks[1], marks[2]
    step = evs[lo:hi]
    t0 = step[0]["ts"]
    streams = sorted({e["args"].get("stream", -1) for e in step})
    last_end = {}
    print(f"# one step: {len(step)} device activities on streams {streams}; span {step[-1]['ts'] + step[-1]['dur'] - t0:.1f} us")
    print("#  start_us   dur_us  gap_us  stream  name")
    for e in step:
        s = e["args"].get("stream", -1)
        gap = e["ts"] - last_end.get(s, e["ts"])
        last_end[s] = e["ts"] + e["dur"]
        if e["dur"] >= args.min_us:
            name = e["name"].replace("upnerf::(anonymous namespace)::", "").replace("void ", "")
            print(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {gap:7.1f}  {streams.index(s):3d}    {name[:70]}")


if __name__ == "__main__":
    main()
