"""Host-side logic of ray-sharded data parallelism, on CPU with gloo, world size 2.

The CUDA kernels cannot run here, so the per-rank "render" is a small differentiable stand-in;
what is under test is the framework's own N>1 plumbing: parameters re-homed in ONE flat buffer
(`FlatGroup`), one mean all-reduce per flat gradient buffer (`allreduce_mean_`, DDP semantics of
the reference's `Trainer(strategy="ddp")`, train.py:72), phase-dead parameters contributing
zeros consistently, and identical Adam steps on every rank with no broadcast.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_params(seed):
    g = torch.Generator().manual_seed(seed)
    live = nn.Parameter(torch.randn(7, 5, generator=g))
    table = nn.Parameter(torch.randn(11, 3, generator=g))      # an embedding table (dense grads)
    dead = nn.Parameter(torch.randn(4, generator=g))           # a head with no gradient in this phase
    return [live, table, dead]


def _local_loss(params, x, idx):
    live, table, _ = params
    return ((x @ live.t()).tanh().sum(1) * table[idx].sum(1)).mean()


def _batch(world_rays, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(world_rays, 5, generator=g), torch.randint(0, 11, (world_rays,), generator=g)


def _worker(rank, world, port, out_dir):
    from upnerf_b200.models.nerf_system import FlatGroup, allreduce_mean_

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _make_params(seed=1)                           # every rank seeds identically (train.py:15-20)
        grp = FlatGroup(params, "cpu")
        opt = torch.optim.Adam([grp.flat], lr=1e-2, eps=1e-8)
        x, idx = _batch(8 * world)
        lo, hi = rank * 8, (rank + 1) * 8                       # rays shard; weights replicate
        for _ in range(3):
            grp.zero_grad()
            # gradients land in the flat buffer through the views (what the kernels do on the GPU)
            loss = _local_loss(grp.params, x[lo:hi], idx[lo:hi])
            gl, gt = torch.autograd.grad(loss, [grp.params[0], grp.params[1]])
            grp.params[0].grad.add_(gl)
            grp.params[1].grad.add_(gt)
            allreduce_mean_(grp.flat.grad)
            opt.step()
        torch.save({"flat": grp.flat.data.clone(), "grad": grp.flat.grad.clone()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_matches_single_process_on_concatenated_batch(tmp_path):
    from upnerf_b200.models.nerf_system import FlatGroup

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    # replicas stay bit-identical without any broadcast
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])
    assert torch.equal(outs[0]["grad"], outs[1]["grad"])

    # single process on the concatenated batch: mean over ranks of per-rank means == global mean
    params = _make_params(seed=1)
    grp = FlatGroup(params, "cpu")
    opt = torch.optim.Adam([grp.flat], lr=1e-2, eps=1e-8)
    x, idx = _batch(8 * world)
    for _ in range(3):
        grp.zero_grad()
        loss = _local_loss(grp.params, x, idx)
        gl, gt = torch.autograd.grad(loss, [grp.params[0], grp.params[1]])
        grp.params[0].grad.add_(gl)
        grp.params[1].grad.add_(gt)
        opt.step()
    assert torch.allclose(outs[0]["flat"], grp.flat.data, rtol=1e-5, atol=1e-6)
    # the phase-dead tensor got an exactly-zero gradient everywhere and (Adam with g=0, m=v=0) did not move
    off, k = grp.offsets[id(grp.params[2])]
    assert torch.equal(outs[0]["grad"][off:off + k], torch.zeros(k))
    assert torch.equal(outs[0]["flat"][off:off + k], _make_params(seed=1)[2].data)


def test_flat_group_views_alias_the_flat_buffers():
    from upnerf_b200.models.nerf_system import FlatGroup

    params = _make_params(seed=2)
    before = [p.data.clone() for p in params]
    grp = FlatGroup(params, "cpu")
    assert grp.flat.numel() == sum(p.numel() for p in params)
    for p, b in zip(params, before):
        assert torch.equal(p.data, b)                            # values preserved
        off, k = grp.offsets[id(p)]
        assert p.data.data_ptr() == grp.flat.data[off:].data_ptr()
        assert p.grad.data_ptr() == grp.flat.grad[off:].data_ptr()
    grp.flat.grad.fill_(1.0)
    assert all(float(p.grad.sum()) == p.numel() for p in params)
    sl = grp.slice_of(params[:2])()
    assert sl.numel() == params[0].numel() + params[1].numel()


def test_bench_rank_sharding_is_disjoint_and_weak():
    """bench.py gives every rank its own seeded batch of the same size (weak scaling)."""
    import bench

    a = bench.host_batch(64, seed=100, pinned=False)
    b = bench.host_batch(64, seed=101, pinned=False)
    assert a["directions"].shape == b["directions"].shape == (64, 3)
    assert not torch.equal(a["img_idx"], b["img_idx"])
