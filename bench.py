#!/usr/bin/env python
"""Benchmark of the UP-NeRF train hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full optimisation step of BASELINE.json config 2/3 on one batch of synthetic
rays per GPU: SE(3) pose refinement + ray casting, coarse (64) + fine (64+64) NeRF-W passes with
sample_pdf, appearance/candidate embeddings, compositing, TransientNet (beta uncertainty) +
UPNeRFLoss, backward to every parameter / embedding / pose, gradient all-reduce (N > 1) and both
Adam updates.  `value` = rays/s with the batch resident in HBM; `e2e` = the same step fed from
pinned HOST buffers (H2D inside the timed region) with the loss read back (D2H).

`--impl reference` times the reference's CPU path instead: the oracle port of the same step
(oracle/upnerf_oracle.py -- the Python reference itself cannot travel to the GPU box) on all host
threads, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

N_IMG = 763                      # brandenburg_gate train images
S_C, N_IMP = 64, 64
PROGRESS = 0.30                  # phase 1 (sched_mult = 0.5): the superset of work, headline phase
MACS = {0: 755_584, 1: 814_720, 2: 714_240}     # per-sample MACs of the reference MLP
FLOP_PER_RAY = (S_C + S_C + N_IMP) * MACS[1] * 2 * 3
CATS = ["gemm_tc", "wgrad_tc", "gemm_simt", "composite", "posenc", "sampling", "pose_rays", "heads", "pack",
        "trunk_fwd", "trunk_bwd"]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def host_batch(R, seed, pinned):
    """One synthetic training batch in the reference's layout (datasets/phototourism.py:420-454,
    SURVEY.md 8(a0)): random pixels of 512x384 / f=400 cameras, random image ids, identity poses
    (pose.noise = -1), U[0,1) colours, unit-norm 384-d features, inverse depths in [1/far, 1/near]."""
    g = torch.Generator().manual_seed(seed)
    near, far = 0.1, 5.0
    px = torch.randint(0, 512, (R,), generator=g).float()
    py = torch.randint(0, 384, (R,), generator=g).float()
    feats = torch.nn.functional.normalize(torch.randn(R, 384, generator=g), dim=-1)
    b = {
        "ray_infos": torch.tensor([[near, far]]).repeat(R, 1),
        "directions": torch.stack([(px - 256) / 400.0, -(py - 192) / 400.0, -torch.ones(R)], -1),
        "c2w": torch.eye(3, 4).expand(R, 3, 4).contiguous(),
        "rgbs": torch.rand(R, 3, generator=g),
        "feats": feats,
        "img_idx": torch.randint(0, N_IMG, (R,), generator=g),
        "inv_depths": 1 / far + (1 / near - 1 / far) * torch.rand(R, generator=g),
    }
    if pinned:
        b = {k: v.pin_memory() for k, v in b.items()}
    return b


def make_system(precision, device):
    from upnerf_b200.models.nerf_system import NeRFSystem

    hp = {"nerf.N_samples": S_C, "nerf.N_importance": N_IMP, "max_steps": 600000, "kernel.precision": precision}
    torch.manual_seed(42)                                       # reference default seed
    s = NeRFSystem(hp, N_images_train=N_IMG, device=device)     # reference constructors' default init
    s.set_progress(PROGRESS)
    s.global_step = int(PROGRESS * 2 * hp["max_steps"])
    return s


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def cpu_port_rate(R_sample, steps, warmup, threads):
    """rays/s of the oracle port of the same train step on the host cores."""
    from oracle import synth
    from oracle import upnerf_oracle as O
    from oracle.train_step import KW, OracleSystem, rng_for

    torch.set_num_threads(threads)
    torch.set_flush_denormal(True)      # denormal transmittances make phase 0 4x slower otherwise
    cfgs = {"nerf_coarse": O.NerfConfig(typ="coarse", **KW), "nerf_fine": O.NerfConfig(typ="fine", **KW)}
    sd = {}
    for k, cfg in cfgs.items():
        for pn, v in synth.nerf_state(cfg, 11 + (k == "nerf_fine"), gain=1.0).items():
            sd[f"{k}.{pn}"] = v
    for k, v in synth.embeddings(N_IMG, cfgs["nerf_coarse"], 11).items():
        sd[f"embedding_{k}.weight"] = v
    for pn, v in synth.transient_state(N_IMG, 11).items():
        sd[f"transient_net.{pn}"] = v
    sd["se3_refine.weight"] = torch.zeros(N_IMG, 6)
    sd["depth_scale.weight"] = torch.zeros(N_IMG, 2)
    orc = OracleSystem(cfgs, sd, N_IMG, S_C, N_IMP, 600000)
    orc.progress, orc.step_no = PROGRESS, int(PROGRESS * 600000)
    times = []
    for it in range(warmup + steps):
        b = host_batch(R_sample, 1000 + it, pinned=False)
        rng = rng_for(R_sample, S_C, N_IMP, O.schedule_mult(PROGRESS), 2000 + 10 * it)
        orc.progress = PROGRESS
        t0 = time.perf_counter()
        orc.step(b, rng)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return R_sample / (sum(times) / len(times)), sum(times) / len(times)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the step (oracle port: the Python
    reference cannot travel to the GPU box), all host threads, bounded number of full batches."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    R_sample = args.rays
    n_steps, n_warm = max(1, min(args.steps, 3)), 1
    rate, sec = cpu_port_rate(R_sample, n_steps, n_warm, threads)
    sample = (f"{n_steps} timed + {n_warm} warm-up steps of {R_sample} rays of the same train step (64+64 samples, phase 1, "
              f"{N_IMG} images), oracle port, {threads} torch threads, flush-denormal on")
    line = {
        "impl": "reference", "metric": "train rays/s (fwd+bwd, 64+64 samp/ray)", "value": rate, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "fp32"),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(kernel):
    """DRAM bytes measured by `ncu --set full` for `kernel` (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py from the committed captures): (sum dram bytes, sum algorithmic bytes, source)."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text()).get(kernel)
    if not d:
        return None
    return sum(x["dram_bytes"] for x in d["launches"]), sum(x["algorithmic_bytes"] for x in d["launches"]), d["source"]


def workload_config(args, precision):
    return {"workload": "BASELINE config 2+3: full UP-NeRF train step (pose refinement + coarse/fine render with "
                        "sample_pdf + embeddings + TransientNet/beta loss + backward + 2x Adam), phase 1 (sched_mult 0.5)",
            "rays_per_gpu": args.rays, "n_samples": S_C, "n_importance": N_IMP, "n_images": N_IMG,
            "precision": precision, "parallelism": f"ray-sharded dp{args.gpus}",
            "l2_policy": "inputs larger than L2: every step streams >5 GB of activations through the 126 MB L2; "
                         "no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from upnerf_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    R = args.rays
    system = make_system(args.precision, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM -------------------------------------------------------
    n_batches = 4
    dev_batches = [{k: v.to(dev) for k, v in host_batch(R, 10 * rank + i, False).items()} for i in range(n_batches)]
    for i in range(W):
        system.training_step(dev_batches[i % n_batches], i)
    barrier()
    lib = L.lib()
    lib.upnerf_launch_count.restype = __import__("ctypes").c_longlong
    launches0 = lib.upnerf_launch_count()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        system.training_step(dev_batches[i % n_batches], i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    launches = (lib.upnerf_launch_count() - launches0)
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * R / (ms * 1e-3)

    # ---- leg 2: end to end from pinned host buffers -------------------------------------------
    pinned = [host_batch(R, 100 + 10 * rank + i, True) for i in range(n_batches)]
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    from upnerf_b200.utils.pipeline import DelayedScalar, DevicePrefetcher

    def e2e_run(n, first):
        """n steps fed from pinned host batches: the H2D copy of batch i+1 rides a copy stream under
        step i, the loss of step i-1 is read on the host while step i runs (one D2H read per step; the
        last one is drained before the region ends) -- all of it inside the timed region."""
        losses, tail = [], DelayedScalar()
        feed = DevicePrefetcher((pinned[(first + i) % n_batches] for i in range(n)), dev)
        for i, b in enumerate(feed):
            v = tail.push(system.training_step(b, first + i))
            if v is not None:
                losses.append(v)
        losses.append(tail.last())
        return losses

    e2e_run(3, 0)
    barrier()
    e0.record()
    e2e_losses = e2e_run(K, 3)
    e1.record()
    barrier()
    assert len(e2e_losses) == K and all(math.isfinite(x) for x in e2e_losses), "e2e: a step's loss was not read"
    ms_e2e = e0.elapsed_time(e1) / K
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t)

    # ---- per-kernel-family device time (CUDA events on the launching stream) ------------------
    import ctypes as C

    # (every rank runs these steps: training_step contains the gradient all-reduce, a collective)
    prof = {}
    lib.upnerf_profile_enable(1)
    n_prof = min(K, 5)
    for i in range(n_prof):
        system.training_step(dev_batches[i % n_batches], i)
    torch.cuda.synchronize()
    n = len(CATS)
    ms_a, ln_a, wk_a, by_a = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.check(lib.upnerf_profile_collect(ms_a, ln_a, wk_a, by_a, n), "upnerf_profile_collect")
    lib.upnerf_profile_enable(0)
    for i, c in enumerate(CATS):
        prof[c] = {"ms_per_step": ms_a[i] / n_prof, "launches_per_step": ln_a[i] / n_prof,
                   "work_per_step": wk_a[i] / n_prof, "bytes_per_step": by_a[i] / n_prof}
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    fam_total = sum(p["ms_per_step"] for p in prof.values())

    def hbm_view(cat, kernel):
        p = prof[cat]
        n = max(p["launches_per_step"], 1)
        gbs = p["bytes_per_step"] / (p["ms_per_step"] * 1e-3) / 1e9 if p["ms_per_step"] > 0 else 0.0
        r = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
             "frac": gbs / pk["hbm_gbs"], "peak_source": f"{pk['source']} hbm_gbs (copy bandwidth)",
             "algorithmic_bytes_per_launch": p["bytes_per_step"] / n, "ms_per_launch": p["ms_per_step"] / n,
             "launches_per_step": n, "share_of_step": p["ms_per_step"] / fam_total, "traffic": None}
        t = ncu_traffic(kernel)
        if t:   # measured DRAM bytes scaled from the profiled launches to this run's average launch
            r["traffic"] = r["algorithmic_bytes_per_launch"] * t[0] / t[1]
            r["traffic_source"] = t[2]
        return r

    def tensor_view(cats, kernel):
        ms_ = sum(prof[c]["ms_per_step"] for c in cats)
        fl = sum(prof[c]["work_per_step"] for c in cats)
        n = max(sum(prof[c]["launches_per_step"] for c in cats), 1)
        tf = fl / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
        return {"bound": "tensor", "kernel": kernel, "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": tf / pk["tf_sustained"],
                "peak_source": f"{pk['source']} bf16_tflops_sustained (kernels timed inside a long step)",
                "flop_per_launch": fl / n, "ms_per_launch": ms_ / n, "share_of_step": ms_ / fam_total}

    # The kernel with the largest share of the step is the weight-gradient GEMM; it streams both of
    # its operands from HBM once (128 flop/byte at N = K = 256, below the machine balance of ~220),
    # so HBM bandwidth bounds it.  The tensor-pipe view of the fused MLP kernels and of all tcgen05
    # kernels together follows as extra objects.
    roofline = hbm_view("wgrad_tc", "wgrad_tc_kernel")
    all_tc = ["gemm_tc", "wgrad_tc", "trunk_fwd", "trunk_bwd"]
    roofline_mlp = {
        "fused_trunk": tensor_view(["trunk_fwd", "trunk_bwd"], "mlp_trunk_fwd_kernel + mlp_trunk_bwd_kernel"),
        "fused_trunk_hbm": {c: hbm_view(c, k) for c, k in (("trunk_fwd", "mlp_trunk_fwd_kernel"),
                                                            ("trunk_bwd", "mlp_trunk_bwd_kernel"))},
        "all_tcgen05": tensor_view(all_tc, "all tcgen05 kernels (fused trunk, layer GEMMs, weight gradients)"),
        "whole_step_algorithmic": {"flop_per_ray": FLOP_PER_RAY, "tflops": value / world * FLOP_PER_RAY / 1e12,
                                   "frac_of_sustained_peak": value / world * FLOP_PER_RAY / 1e12 / pk["tf_sustained"]},
        "families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in prof.items()},
    }
    line = {
        "metric": "train rays/s (fwd+bwd, 64+64 samp/ray)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args, args.precision),
        "e2e": {"value": world * R / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "pipeline": "public API: NeRFSystem.training_step fed by utils.pipeline.DevicePrefetcher (H2D of "
                            "batch i+1 from pinned memory on a copy stream under step i) + DelayedScalar (loss of "
                            "step i-1 read on the host during step i; last one drained inside the timed region)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "roofline_mlp": roofline_mlp,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = cpu_port_rate(2048, 4, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"2048 rays/step x 4 timed steps (+1 warm-up) of the same train step (oracle port, "
                                          f"{threads} torch threads, flush-denormal on), {sec:.2f} s/step"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
