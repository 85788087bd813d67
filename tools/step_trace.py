"""Timeline of the train step from CUPTI (torch.profiler): device busy time vs idle gaps inside the
step, and device time per kernel name (our kernels AND the torch ones around them).

    python tools/step_trace.py [--steps 6] [--rays 4096] [--top 45]

Numbers taken under the profiler are diagnostic only (never a bench value).
"""
from __future__ import annotations

import argparse
import collections
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    system = bench.make_system("bf16", dev)
    batches = [{k: v.to(dev) for k, v in bench.host_batch(args.rays, i, False).items()} for i in range(4)]
    for i in range(5):
        system.training_step(batches[i % 4], i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            system.training_step(batches[i % 4], i)
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            evs.append((e.time_range.start, e.time_range.end, e.name))
    evs.sort()
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    busy, cur_end, gaps = 0.0, evs[0][0], []
    for s, e, n in evs:
        if s > cur_end:
            gaps.append((s - cur_end, n))
            cur_end = s
        if e > cur_end:
            busy += e - cur_end
            cur_end = e
    span = t1 - t0
    by = collections.defaultdict(lambda: [0.0, 0])
    for s, e, n in evs:
        by[n][0] += e - s
        by[n][1] += 1
    rows = sorted(by.items(), key=lambda kv: -kv[1][0])
    out = {"steps": args.steps, "span_ms_per_step": span / 1e3 / args.steps, "busy_ms_per_step": busy / 1e3 / args.steps,
           "idle_ms_per_step": (span - busy) / 1e3 / args.steps, "kernels_per_step": len(evs) / args.steps}
    print(json.dumps(out))
    ours = sum(v[0] for k, v in by.items() if "upnerf" in k) / 1e3 / args.steps
    print(f"our kernels {ours:.3f} ms/step, others {busy / 1e3 / args.steps - ours:.3f} ms/step")
    for name, (us, n) in rows[: args.top]:
        print(f"{us / 1e3 / args.steps:8.4f} ms/step {n / args.steps:7.1f} x {us / n:8.1f} us  {name[:110]}")
    gaps.sort(reverse=True)
    print("largest gaps (us, next kernel):")
    for g, n in gaps[:15]:
        print(f"  {g:8.1f}  {n[:90]}")
    gap_by = collections.defaultdict(float)
    for g, n in gaps:
        gap_by[n] += g
    print("gap time by following kernel (ms/step):")
    for n, g in sorted(gap_by.items(), key=lambda kv: -kv[1])[:15]:
        print(f"  {g / 1e3 / args.steps:8.4f}  {n[:90]}")
    if args.out:
        Path(args.out).write_text(json.dumps({"summary": out, "kernels": [(k, v[0] / args.steps, v[1] / args.steps) for k, v in rows]}))


if __name__ == "__main__":
    main()
