"""Host-side cost of the train step: wall time the Python/C host spends ISSUING one step (no device
sync inside the loop) against the device time of the same steps, plus per-section host time.

    python tools/host_time.py [--steps 40] [--rays 4096]

If issue time per step is close to device time per step the step is host-bound in places (idle gaps).
"""
from __future__ import annotations

import argparse
import collections
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--rays", type=int, default=4096)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    system = bench.make_system("bf16", dev)
    batches = [{k: v.to(dev) for k, v in bench.host_batch(args.rays, i, False).items()} for i in range(4)]
    for i in range(5):
        system.training_step(batches[i % 4], i)
    torch.cuda.synchronize()

    acc = collections.defaultdict(float)

    def wrap(obj, name, label):
        f = getattr(obj, name)

        def g(*a, **k):
            t = time.perf_counter()
            r = f(*a, **k)
            acc[label] += time.perf_counter() - t
            return r
        setattr(obj, name, g)

    from upnerf_b200 import _lib as L
    from upnerf_b200.models import nerf_system as NS
    from upnerf_b200.models import rendering as RR
    from upnerf_b200.utils import ray as RU
    wrap(L, "render_fwd", "C render_fwd")
    wrap(L, "render_bwd", "C render_bwd")
    wrap(RR, "render_rays", "render_rays (python + C fwd)")
    wrap(RU, "refine_rays", "refine_rays")
    wrap(NS, "fused_tail", "fused_tail")
    wrap(torch.autograd, "backward", "autograd.backward (incl. C bwd)")
    wrap(system.transient_net, "forward", "transient_net fwd")
    for i, (o, s) in enumerate(zip(system._optimizers, system._schedulers)):
        wrap(o, "step", "optimizer.step")
        wrap(s, "step", "scheduler.step")
    wrap(system, "set_progress", "set_progress")
    wrap(system.group_main, "zero_grad", "zero_grad")
    wrap(system.group_pose, "zero_grad", "zero_grad")

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        system.training_step(batches[i % 4], i)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"issue {t_issue / args.steps * 1e3:.3f} ms/step, device {e0.elapsed_time(e1) / args.steps:.3f} ms/step, "
          f"wall {t_all / args.steps * 1e3:.3f} ms/step")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {v / args.steps * 1e3:8.3f} ms/step  {k}")
    # the same loop with a sync per step = pure host issue latency when the queue is empty
    acc.clear()
    t_host = 0.0
    for i in range(args.steps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        system.training_step(batches[i % 4], i)
        t_host += time.perf_counter() - t
    torch.cuda.synchronize()
    print(f"issue with empty queue: {t_host / args.steps * 1e3:.3f} ms/step")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {v / args.steps * 1e3:8.3f} ms/step  {k}")


if __name__ == "__main__":
    main()
