"""GPU, world size 2 over NCCL (skipped with fewer than two devices): the gradients `training_step`
leaves in the flat buffers after its all-reduce must equal the single-process gradients of the
CONCATENATED batch (DDP semantics of the reference, train.py:72; SURVEY.md section 4 tier 3), and the
replicas must stay bit-identical after their Adam steps without any broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth
from oracle import upnerf_oracle as O
from oracle.train_step import rng_for

pytestmark = pytest.mark.gpu
R, S, NI, N_IMG, MAX_STEPS, START = 256, 32, 32, 12, 1000, 0.3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _system(device, precision):
    from test_train_step_gpu import make_system

    sys_, _, _ = make_system(N_IMG, S, NI, precision, MAX_STEPS, device)
    sys_.set_progress(START)
    sys_.global_step = int(round(START * 2 * MAX_STEPS))
    return sys_


def _batch_and_rng(lo, hi, device):
    b = synth.ray_batch(R, N_IMG, 100)
    m = O.schedule_mult(float(torch.tensor(START)))
    rng = rng_for(R, S, NI, m, 200)
    b = {k: v[lo:hi].contiguous().to(device) for k, v in b.items()}
    rng = dict(perturb_rand=rng["perturb_rand"][lo:hi].contiguous(), u=[u[lo:hi].contiguous() for u in rng["u"]])
    return b, rng


def _worker(rank, world, port, out_dir, precision):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sys_ = _system(dev, precision)
        n = R // world
        b, rng = _batch_and_rng(rank * n, (rank + 1) * n, dev)
        loss = sys_.training_step(b, 0, rng=rng)
        torch.cuda.synchronize()
        torch.save({"g_main": sys_.group_main.flat.grad.cpu(), "g_pose": sys_.group_pose.flat.grad.cpu(),
                    "p_main": sys_.group_main.flat.data.cpu(), "p_pose": sys_.group_pose.flat.data.cpu(),
                    "loss": float(loss)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_gradients_equal_concatenated_batch(cuda_dev, tmp_path, precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), precision), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for k in ("g_main", "g_pose", "p_main", "p_pose"):          # replicas identical, no broadcast needed
        assert torch.equal(outs[0][k], outs[1][k]), k
    sys_ = _system(cuda_dev, precision)
    b, rng = _batch_and_rng(0, R, cuda_dev)
    loss = sys_.training_step(b, 0, rng=rng)
    # the loss terms are means over the local rays: mean over ranks of local means == global mean
    assert abs(0.5 * (outs[0]["loss"] + outs[1]["loss"]) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    tol = 1e-5 if precision == "fp32" else 1e-3      # bf16: tile boundaries fall differently in a half batch
    for k, flat in (("g_main", sys_.group_main.flat.grad), ("g_pose", sys_.group_pose.flat.grad)):
        ref = flat.cpu().double()
        err = float((outs[0][k].double() - ref).norm() / ref.norm())
        print(f"[{precision}] {k}: 2-rank mean vs concatenated batch, relative {err:.2e}")
        assert err <= tol, (k, err)


def _worker_steps(rank, world, port, out_dir, graphed, n_steps):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sys_ = _system(dev, "bf16")
        sys_.hparams["kernel.cuda_graph"] = graphed
        sys_.hparams["nerf.perturb"] = 0.0          # deterministic sampling: the runs are comparable step by step
        n = R // world
        losses = []
        for it in range(n_steps):
            b = synth.ray_batch(R, N_IMG, 500 + it)
            b = {k: v[rank * n:(rank + 1) * n].contiguous().to(dev) for k, v in b.items()}
            losses.append(float(sys_.training_step(b, it)))
        torch.cuda.synchronize()
        torch.save({"p_main": sys_.group_main.flat.data.cpu(), "p_pose": sys_.group_pose.flat.data.cpu(),
                    "losses": losses, "replays": sys_.graph_replays},
                   os.path.join(out_dir, f"{'g' if graphed else 'e'}{rank}.pt"))
        sys_.release_graphs()       # before the communicator goes away (see NeRFSystem.release_graphs)
    finally:
        dist.destroy_process_group()


def test_two_rank_graphed_steps(cuda_dev, tmp_path):
    """The step as a CUDA-graph replay WITH its two NCCL all-reduces captured (the fine network's slice goes out
    between the two backward passes): replicas stay bit-identical without a broadcast, and the losses follow the
    eagerly launched 2-rank run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, n_steps = 2, 8
    for graphed in (False, True):
        mp.spawn(_worker_steps, args=(world, _free_port(), str(tmp_path), graphed, n_steps), nprocs=world, join=True)
    e = [torch.load(tmp_path / f"e{r}.pt") for r in range(world)]
    g = [torch.load(tmp_path / f"g{r}.pt") for r in range(world)]
    assert g[0]["replays"] == n_steps - 3 and e[0]["replays"] == 0
    for k in ("p_main", "p_pose"):
        assert torch.equal(g[0][k], g[1][k]), k              # graphed replicas identical
        assert torch.equal(e[0][k], e[1][k]), k
    for r in range(world):
        for i, (a, b) in enumerate(zip(e[r]["losses"], g[r]["losses"])):
            assert abs(a - b) <= 2e-2 * max(1.0, abs(a)), (r, i, a, b)
