#!/usr/bin/env python
"""Benchmark of the UP-NeRF train hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload train|render]

A "step" is one full optimisation step of BASELINE.json config 2/3 on one batch of synthetic
rays per GPU: SE(3) pose refinement + ray casting, coarse (64) + fine (64+64) NeRF-W passes with
sample_pdf, appearance/candidate embeddings, compositing, TransientNet (beta uncertainty) +
UPNeRFLoss, backward to every parameter / embedding / pose, gradient all-reduce (N > 1) and both
Adam updates.  `value` = rays/s with the batch resident in HBM; `e2e` = the same step fed from
pinned HOST buffers (H2D inside the timed region) with the loss read back (D2H).

Rays per GPU default to BASELINE config 2/3 (4096) on one GPU and to config 4 (8192 per GPU) for N > 1.

`--impl reference` times the reference's CPU path instead: the oracle port of the same step
(oracle/upnerf_oracle.py -- the Python reference itself cannot travel to the GPU box) on all host
threads; it runs exactly the --steps / --warmup it prints, each step on a bounded sample of the batch.

`--workload render` measures BASELINE config 5 instead: chunked no-grad render of one 1920x1080 image
(64+64 samples, phase 2, val.chunk_size 4096 as the reference), one image per step.
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
        "trunk_fwd", "trunk_bwd", "wgrad_reduce", "tnet", "gemm_tf32"]
TRAIN_METRIC = "train rays/s (fwd+bwd, 64+64 samp/ray)"
RENDER_METRIC = "inference render rays/s (1920x1080, 64+64 samp/ray, no grad)"
RENDER_W, RENDER_H, RENDER_CHUNK = 1920, 1080, 4096     # BASELINE config 5; val.chunk_size of configs/default.yaml
FLOP_PER_RAY_RENDER = (S_C + S_C + N_IMP) * MACS[2] * 2


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


def cpu_port_rate(R_sample, steps, warmup, threads, device="cpu"):
    """rays/s of the oracle port of the same train step on the host cores (device="cpu") or, as the
    "reference CUDA-eager" number, the same plain-PyTorch port with its tensors on the GPU."""
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
    dev = torch.device(device)
    sd = {k: v.to(dev) for k, v in sd.items()}
    orc = OracleSystem(cfgs, sd, N_IMG, S_C, N_IMP, 600000)
    orc.progress, orc.step_no = PROGRESS, int(PROGRESS * 600000)
    times = []
    for it in range(warmup + steps):
        b = {k: v.to(dev) for k, v in host_batch(R_sample, 1000 + it % 4, pinned=False).items()}
        rng = rng_for(R_sample, S_C, N_IMP, O.schedule_mult(PROGRESS), 2000 + 10 * (it % 4))
        rng = dict(perturb_rand=rng["perturb_rand"].to(dev), u=[u.to(dev) for u in rng["u"]])
        orc.progress = PROGRESS
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss, _ = orc.step(b, rng)
        if dev.type == "cuda":
            float(loss)                      # the step's loss read on the host, like a training loop's log
            torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return R_sample / (sum(times) / len(times)), sum(times) / len(times)


def cpu_render_rate(R_sample, steps, warmup, threads):
    """rays/s of the oracle port of the config-5 render (no grad, phase 2, perturb 0) on the host cores."""
    from oracle import synth
    from oracle import upnerf_oracle as O
    from oracle.train_step import KW

    torch.set_num_threads(threads)
    torch.set_flush_denormal(True)
    cfgs = {"nerf_coarse": O.NerfConfig(typ="coarse", **KW), "nerf_fine": O.NerfConfig(typ="fine", **KW)}
    sds = {k: synth.nerf_state(cfg, 11 + (k == "nerf_fine"), progress=0.75, gain=1.0) for k, cfg in cfgs.items()}
    emb = synth.embeddings(N_IMG, cfgs["nerf_coarse"], 11)
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            b = host_batch(R_sample, 1000 + it % 4, pinned=False)
            o, d = O.get_rays(b["directions"], b["c2w"])
            rays = torch.cat([o, d, b["ray_infos"]], 1)
            t0 = time.perf_counter()
            O.render_rays(sds, cfgs, emb, rays, b["img_idx"], 1.0, 0.75, N_samples=S_C, perturb=0.0, N_importance=N_IMP)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return R_sample / (sum(times) / len(times)), sum(times) / len(times)


def bounded_sample(rays, steps, warmup, rays_per_s_guess, budget_s=150.0):
    """Rays per CPU step such that steps + warmup steps take about `budget_s` seconds (the rate in rays/s
    does not depend on the batch size on the CPU), at most the workload's own batch, a multiple of 256."""
    r = int(budget_s * rays_per_s_guess / max(1, steps + warmup))
    return max(256, min(rays, r // 256 * 256))


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the step (oracle port: the Python
    reference cannot travel to the GPU box), all host threads.  Runs exactly --steps timed and --warmup
    untimed steps; each step processes a bounded sample of the batch (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_steps, n_warm = max(1, args.steps), max(0, args.warmup)
    if args.workload == "render":
        R_sample = bounded_sample(RENDER_CHUNK, n_steps, n_warm, 2500.0)
        rate, sec = cpu_render_rate(R_sample, n_steps, n_warm, threads)
        what = f"the config-5 render (no grad, phase 2, 64+64 samples, {N_IMG} images)"
        metric, cfg = RENDER_METRIC, render_config(args, "fp32")
    else:
        R_sample = bounded_sample(args.rays, n_steps, n_warm, 850.0)
        rate, sec = cpu_port_rate(R_sample, n_steps, n_warm, threads)
        what = f"the same train step (64+64 samples, phase 1, {N_IMG} images)"
        metric, cfg = TRAIN_METRIC, workload_config(args, "fp32")
    sample = (f"{n_steps} timed + {n_warm} warm-up steps of {R_sample} rays each of {what}, oracle port, "
              f"{threads} torch threads, flush-denormal on")
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": n_steps, "warmup": n_warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
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
    which = "2+3" if args.gpus == 1 else "4 (= config 3 ray-sharded, gradient all-reduce over NCCL)"
    return {"workload": f"BASELINE config {which}: full UP-NeRF train step (pose refinement + coarse/fine render "
                        "with sample_pdf + embeddings + TransientNet/beta loss + backward + 2x Adam), phase 1 "
                        f"(sched_mult 0.5), {args.rays} rays per GPU",
            "rays_per_gpu": args.rays, "n_samples": S_C, "n_importance": N_IMP, "n_images": N_IMG,
            "precision": precision, "parallelism": f"ray-sharded dp{args.gpus}",
            "l2_policy": "inputs larger than L2: every step streams >5 GB of activations through the 126 MB L2; "
                         "no explicit flush"}


def render_config(args, precision):
    return {"workload": f"BASELINE config 5: chunked full-image inference render {RENDER_W}x{RENDER_H} = "
                        f"{RENDER_W * RENDER_H} rays per step, coarse+fine 64+64 samples, phase 2 (sched_mult 1), perturb 0, "
                        f"torch.no_grad, val.chunk_size {RENDER_CHUNK} (NeRFSystem.forward(train=False), "
                        "models/nerf_system.py:93-148); image rows shard across GPUs",
            "rays_per_step": RENDER_W * RENDER_H, "chunk": RENDER_CHUNK, "n_samples": S_C, "n_importance": N_IMP,
            "n_images": N_IMG, "precision": precision, "parallelism": f"row-sharded x{args.gpus}",
            "l2_policy": "inputs larger than L2: each image streams >100 GB of activations; no explicit flush"}


def dist_setup():
    import torch.distributed as dist

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    return dist, world, rank, local, dev, barrier, max_over_ranks


def collect_profile(lib, L, n_steps):
    import ctypes as C

    n = len(CATS)
    ms_a, ln_a, wk_a, by_a = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.check(lib.upnerf_profile_collect(ms_a, ln_a, wk_a, by_a, n), "upnerf_profile_collect")
    return {c: {"ms_per_step": ms_a[i] / n_steps, "launches_per_step": ln_a[i] / n_steps,
                "work_per_step": wk_a[i] / n_steps, "bytes_per_step": by_a[i] / n_steps} for i, c in enumerate(CATS)}


def roofline_views(prof, pk):
    fam_total = sum(p["ms_per_step"] for p in prof.values())

    def hbm_view(cat, kernel):
        p = prof[cat]
        n = max(p["launches_per_step"], 1)
        gbs = p["bytes_per_step"] / (p["ms_per_step"] * 1e-3) / 1e9 if p["ms_per_step"] > 0 else 0.0
        r = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
             "frac": gbs / pk["hbm_gbs"], "peak_source": f"{pk['source']} hbm_gbs (copy bandwidth)",
             "algorithmic_bytes_per_launch": p["bytes_per_step"] / n, "ms_per_launch": p["ms_per_step"] / n,
             "launches_per_step": n, "share_of_step": p["ms_per_step"] / fam_total if fam_total else 0.0, "traffic": None}
        t = ncu_traffic(kernel)
        if t:   # measured DRAM bytes scaled from the profiled launches to this run's average launch
            r["traffic"] = r["algorithmic_bytes_per_launch"] * t[0] / t[1]
            r["traffic_source"] = t[2]
        return r

    def tensor_view(cats, kernel, peak_key="tf_sustained"):
        ms_ = sum(prof[c]["ms_per_step"] for c in cats)
        fl = sum(prof[c]["work_per_step"] for c in cats)
        n = max(sum(prof[c]["launches_per_step"] for c in cats), 1)
        tf = fl / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
        return {"bound": "tensor", "kernel": kernel, "achieved": tf, "peak": pk[peak_key], "unit": "TFLOP/s",
                "frac": tf / pk[peak_key], "frac_of_burst_peak": tf / pk["tf_burst"],
                "peak_source": f"{pk['source']} bf16_tflops_sustained (kernels timed inside a long step)",
                "flop_per_launch": fl / n, "ms_per_launch": ms_ / n,
                "share_of_step": ms_ / fam_total if fam_total else 0.0, "traffic": None}

    return hbm_view, tensor_view


def run_train(args):
    from upnerf_b200 import _lib as L

    dist, world, rank, local, dev, barrier, max_over_ranks = dist_setup()
    W = max(3, args.warmup)
    K = args.steps
    R = args.rays
    system = make_system(args.precision, dev)

    # ---- leg 1: inputs resident in HBM -------------------------------------------------------
    n_batches = 4
    dev_batches = [{k: v.to(dev) for k, v in host_batch(R, 10 * rank + i, False).items()} for i in range(n_batches)]
    for i in range(W):
        system.training_step(dev_batches[i % n_batches], i)
    barrier()
    lib = L.lib()
    launches0 = L.launch_count()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        system.training_step(dev_batches[i % n_batches], i)
    e1.record()
    barrier()
    launches = L.launch_count() - launches0
    clk = clocks.stop() if clocks else None
    ms = max_over_ranks(e0.elapsed_time(e1) / K)
    value = world * R / (ms * 1e-3)

    # ---- leg 2: end to end from pinned host buffers -------------------------------------------
    pinned = [host_batch(R, 100 + 10 * rank + i, True) for i in range(n_batches)]
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    from upnerf_b200.utils.pipeline import DelayedScalar, DevicePrefetcher

    def e2e_run(n, first):
        """n steps fed from pinned host batches: the H2D copy of batch i+1 rides a copy stream under
        step i, the loss of step i-3 is read on the host while step i runs (one D2H read per step; the
        last ones are drained before the region ends) -- all of it inside the timed region."""
        losses, tail = [], DelayedScalar(depth=3)
        feed = DevicePrefetcher((pinned[(first + i) % n_batches] for i in range(n)), dev)
        for i, b in enumerate(feed):
            v = tail.push(system.training_step(b, first + i))
            if v is not None:
                losses.append(v)
        losses.extend(tail.drain())
        return losses

    e2e_run(W, 0)
    e2e_ms = []
    for rep in range(3):         # three back-to-back regions of K steps; the median is reported (a single
        barrier()                # host hiccup -- page faults, a scheduler tick -- otherwise doubles a 20-step mean)
        e0.record()
        e2e_losses = e2e_run(K, 3 + rep * K)
        e1.record()
        barrier()
        assert len(e2e_losses) == K and all(math.isfinite(x) for x in e2e_losses), "e2e: a step's loss was not read"
        e2e_ms.append(max_over_ranks(e0.elapsed_time(e1) / K))
    ms_e2e = sorted(e2e_ms)[1]

    # ---- per-kernel-family device time (CUDA events on the launching stream) ------------------
    # (every rank runs these steps: training_step contains the gradient all-reduce, a collective)
    # (eager launches: the library's per-launch events cannot be recorded inside a replayed graph)
    graphed = system.hparams["kernel.cuda_graph"]
    system.hparams["kernel.cuda_graph"] = False
    lib.upnerf_profile_enable(1)
    n_prof = min(K, 5)
    for i in range(n_prof):
        system.training_step(dev_batches[i % n_batches], i)
    torch.cuda.synchronize()
    prof = collect_profile(lib, L, n_prof)
    lib.upnerf_profile_enable(0)
    system.hparams["kernel.cuda_graph"] = graphed

    # ---- N > 1: the same step without its gradient all-reduce (communication cost, measured) -----
    ms_nocomm = None
    if world > 1:
        system.hparams["kernel.skip_allreduce"] = True
        for i in range(3):
            system.training_step(dev_batches[i % n_batches], i)
        barrier()
        e0.record()
        for i in range(K):
            system.training_step(dev_batches[i % n_batches], i)
        e1.record()
        barrier()
        ms_nocomm = max_over_ranks(e0.elapsed_time(e1) / K)
        system.hparams["kernel.skip_allreduce"] = False
        dist.barrier()

    if rank != 0:
        if world > 1:
            system.release_graphs()
            dist.destroy_process_group()
        return
    pk = peaks()
    hbm_view, tensor_view = roofline_views(prof, pk)

    # The kernel with the largest share of the step is the weight-gradient GEMM; it streams both of
    # its operands from HBM once (128 flop/byte at N = K = 256, below the machine balance of ~220),
    # so HBM bandwidth bounds it.  (Its split-partial reduction is a separate kernel and family.)  The
    # tensor-pipe view of the fused MLP kernels and of all tcgen05 kernels together follows as extra objects.
    roofline = hbm_view("wgrad_tc", "wgrad_tc_kernel")
    all_tc = ["gemm_tc", "wgrad_tc", "trunk_fwd", "trunk_bwd", "tnet", "gemm_tf32"]
    roofline_mlp = {
        "fused_trunk": tensor_view(["trunk_fwd", "trunk_bwd"], "pp::mlp_trunk_fwd_pp_kernel + pp::mlp_trunk_bwd_pp_kernel"),
        "fused_trunk_fwd": tensor_view(["trunk_fwd"], "pp::mlp_trunk_fwd_pp_kernel"),
        "fused_trunk_bwd": tensor_view(["trunk_bwd"], "pp::mlp_trunk_bwd_pp_kernel"),
        "fused_trunk_hbm": {c: hbm_view(c, k) for c, k in (("trunk_fwd", "pp::mlp_trunk_fwd_pp_kernel"),
                                                            ("trunk_bwd", "pp::mlp_trunk_bwd_pp_kernel"))},
        "gemm_tc_hbm": hbm_view("gemm_tc", "gemm_tc_kernel"),
        "all_tcgen05": tensor_view(all_tc, "all tcgen05 kernels (fused trunk, layer GEMMs, weight gradients)"),
        "whole_step_algorithmic": {"flop_per_ray": FLOP_PER_RAY, "tflops": value / world * FLOP_PER_RAY / 1e12,
                                   "frac_of_sustained_peak": value / world * FLOP_PER_RAY / 1e12 / pk["tf_sustained"],
                                   "frac_of_burst_peak": value / world * FLOP_PER_RAY / 1e12 / pk["tf_burst"]},
        "families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in prof.items()},
        "families_launches_per_step": {k: v["launches_per_step"] for k, v in prof.items()},
    }
    line = {
        "metric": TRAIN_METRIC, "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args, args.precision),
        "e2e": {"value": world * R / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                "ms_per_step_runs": [round(x, 4) for x in e2e_ms], "statistic": "median of 3 regions of K steps",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "pipeline": "public API: NeRFSystem.training_step fed by utils.pipeline.DevicePrefetcher (H2D of "
                            "batch i+1 from pinned memory into preallocated device buffers on a copy stream under step i) + DelayedScalar (loss of "
                            "step i-3 read on the host during step i; the last ones drained inside the timed region)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "roofline_mlp": roofline_mlp,
    }
    if ms_nocomm is not None:
        line["comm"] = {"ms_per_step_without_allreduce": ms_nocomm, "allreduce_ms_exposed": ms - ms_nocomm,
                        "bytes_per_step": 4 * (system.group_main.flat.numel() + system.group_pose.flat.numel())}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = cpu_port_rate(2048, 4, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"2048 rays/step x 4 timed steps (+1 warm-up) of the same train step (oracle port, "
                                          f"{threads} torch threads, flush-denormal on), {sec:.2f} s/step"}
        # the "before" number on the same GPU: the plain-PyTorch port of the reference step, eager, fp32,
        # tensors on the device (the reference itself is eager PyTorch; it cannot travel to this box)
        del system, dev_batches
        torch.cuda.empty_cache()
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            rate_g, sec_g = cpu_port_rate(R, 5, 2, threads, device=str(dev))
            line["cuda_eager_baseline"] = {"value": rate_g, "unit": "rays/s", "kind": "port on cuda (eager PyTorch fp32, TF32 off)",
                                           "ms_per_step": sec_g * 1e3,
                                           "sample": f"{R} rays/step x 5 timed steps (+2 warm-up), same train step, loss read per step"}
        except Exception as e:          # never lose the bench line to the auxiliary baseline
            line["cuda_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        system.release_graphs()
        dist.destroy_process_group()


def run_render(args):
    """BASELINE config 5 through the public API: NeRFSystem.forward(train=False) renders one 1920x1080 image
    per step in val.chunk_size chunks under no_grad.  `value`: rays resident; `e2e`: per step the pixel
    directions of the image come from pinned host memory (H2D) and the rendered rgb image goes back (D2H)."""
    from upnerf_b200 import _lib as L
    from upnerf_b200.utils.ray import get_ray_directions, get_rays

    dist, world, rank, local, dev, barrier, max_over_ranks = dist_setup()
    W, K = max(3, args.warmup), args.steps
    system = make_system(args.precision, dev)
    system.set_progress(0.75)
    system.hparams["val.chunk_size"] = args.chunk
    H, Wd = RENDER_H, RENDER_W
    rows = [r for r in range(H) if r % world == rank] if world > 1 else list(range(H))      # image rows shard
    Kmat = torch.tensor([[0.8 * Wd, 0, Wd / 2], [0, 0.8 * Wd, H / 2], [0, 0, 1]])
    dirs_host = get_ray_directions(H, Wd, Kmat)[rows].reshape(-1, 3).contiguous().pin_memory()
    R = dirs_host.shape[0]
    c2w = torch.eye(3, 4, device=dev)
    idx = torch.zeros(R, dtype=torch.long, device=dev)
    feats = torch.zeros(R, 384, device=dev)
    nf_row = torch.tensor([[0.1, 5.0]], device=dev)
    out_host = torch.empty(R, 3).pin_memory()

    def render(dirs):
        n = dirs.shape[0]
        o, d = get_rays(dirs, c2w)
        rays = torch.cat([o, d, nf_row.expand(n, 2)], 1)
        return system(rays, feats[:n], idx[:n], 1.0, train=False)["rgb_fine"]

    dirs_dev = dirs_host.to(dev)
    with torch.no_grad():
        n_w = min(R, 262144)
        for _ in range(W):               # warm-up on a quarter-megaray slice per step
            render(dirs_dev[:n_w])
        barrier()
        lib = L.lib()
        launches0 = L.launch_count()
        clocks = ClockSampler(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            img = render(dirs_dev)
        e1.record()
        barrier()
        launches = L.launch_count() - launches0
        clk = clocks.stop() if clocks else None
        ms = max_over_ranks(e0.elapsed_time(e1) / K)
        assert img.shape == (R, 3) and bool(torch.isfinite(img).all())
        # e2e
        barrier()
        e0.record()
        for _ in range(K):
            d_in = dirs_host.to(dev, non_blocking=True)
            out_host.copy_(render(d_in), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1) / K)
        assert bool(torch.isfinite(out_host).all())
        lib.upnerf_profile_enable(1)
        render(dirs_dev)
        torch.cuda.synchronize()
        prof = collect_profile(lib, L, 1)
        lib.upnerf_profile_enable(0)
    if rank != 0:
        if world > 1:
            system.release_graphs()
            dist.destroy_process_group()
        return
    pk = peaks()
    hbm_view, tensor_view = roofline_views(prof, pk)
    total_rays = RENDER_W * RENDER_H
    value = total_rays / (ms * 1e-3)
    roofline = tensor_view(["trunk_fwd"], "pp::mlp_trunk_fwd_pp_kernel (forward-only: activations stay on chip)")
    line = {
        "metric": RENDER_METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": render_config(args, args.precision),
        "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": dirs_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                "pipeline": "pinned pixel directions H2D -> get_rays -> NeRFSystem.forward(train=False) in "
                            f"{args.chunk}-ray chunks -> rgb_fine image D2H into pinned memory"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
        "roofline_mlp": {
            "whole_step_algorithmic": {"flop_per_ray": FLOP_PER_RAY_RENDER, "tflops": value / world * FLOP_PER_RAY_RENDER / 1e12,
                                       "frac_of_sustained_peak": value / world * FLOP_PER_RAY_RENDER / 1e12 / pk["tf_sustained"],
                                       "frac_of_burst_peak": value / world * FLOP_PER_RAY_RENDER / 1e12 / pk["tf_burst"]},
            "gemm_tc_hbm": hbm_view("gemm_tc", "gemm_tc_kernel"),
            "families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in prof.items()},
            "families_launches_per_step": {k: v["launches_per_step"] for k, v in prof.items()}},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = cpu_render_rate(4096, 4, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"4096-ray chunk x 4 timed (+1 warm-up) of the same no-grad render (oracle port, "
                                          f"{threads} torch threads, flush-denormal on), {sec:.2f} s/chunk"}
    print(json.dumps(line), flush=True)
    if world > 1:
        system.release_graphs()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render"])
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per train step (default: 4096 on one GPU = "
                    "BASELINE config 2/3, 8192 per GPU on N > 1 = config 4)")
    ap.add_argument("--chunk", type=int, default=RENDER_CHUNK, help="render workload: rays per chunk (val.chunk_size)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.gpus = max(args.gpus, world)
    if args.rays is None:
        args.rays = 4096 if args.gpus == 1 else 8192
    if args.steps is None:
        args.steps = 100 if args.workload == "train" else 10
    if args.impl == "reference":
        return run_reference(args)
    return run_render(args) if args.workload == "render" else run_train(args)


if __name__ == "__main__":
    main()
