"""GPU: TransientNet on the tensor-core path (`upnerf_tnet_fwd/bwd`, reference models/transient_net.py:27-38)
against (1) the golden written by the real reference module (tests/golden/tail.npz) and (2) the same module
evaluated by torch in fp32 with autograd, at the benchmark batch size.  bf16 operands, fp32 accumulation:
outputs within 2e-2 (BASELINE.json), parameter gradients norm-wise."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _net(n_img, seed, dev, precision):
    from upnerf_b200.models.transient_net import TransientNet

    net = TransientNet(N_images=n_img, beta_min=0.1, trasient_dim=128, feat_dim=384)
    net.load_state_dict(synth.transient_state(n_img, seed))
    net = net.to(dev)
    net.precision = precision
    return net


def test_tnet_forward_matches_reference_golden(cuda_dev):
    z = np.load(GOLD / "tail.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    n_img = int(g["img_idx"].max()) + 1
    net = _net(5, 3, cuda_dev, "bf16")
    assert n_img <= 5
    with torch.no_grad():
        out = net(g["feats"].to(cuda_dev), g["img_idx"].to(cuda_dev))
    for k in ("alpha", "rgb", "beta"):
        ref = g[f"t_{k}"]
        assert out[k].shape == ref.shape, (k, out[k].shape, ref.shape)
        err = float((out[k].cpu() - ref).abs().max())
        assert err <= 2e-2 * max(1.0, float(ref.abs().max())), (k, err)
    # and the fp32 torch path of the same module reproduces the golden tightly
    net.precision = "fp32"
    with torch.no_grad():
        out32 = net(g["feats"].to(cuda_dev), g["img_idx"].to(cuda_dev))
    for k in ("alpha", "rgb", "beta"):
        assert float((out32[k].cpu() - g[f"t_{k}"]).abs().max()) <= 1e-5, k


@pytest.mark.parametrize("R", [4096, 1000])
def test_tnet_forward_backward_vs_torch_fp32(cuda_dev, R):
    n_img = 763
    b = synth.ray_batch(R, n_img, 71)
    feats, idx = b["feats"].to(cuda_dev), b["img_idx"].to(cuda_dev)
    cot = {k: synth.uniform(s, 90 + i).to(cuda_dev) for i, (k, s) in enumerate((("alpha", (R, 1)), ("beta", (R, 1)), ("rgb", (R, 3))))}
    res = {}
    for precision in ("fp32", "bf16"):
        net = _net(n_img, 3, cuda_dev, precision)
        out = net(feats, idx)
        sum((out[k] * cot[k]).sum() for k in out).backward()
        res[precision] = ({k: v.detach() for k, v in out.items()}, {k: p.grad.clone() for k, p in net.named_parameters()})
    for k in ("alpha", "rgb", "beta"):
        ref = res["fp32"][0][k]
        err = float((res["bf16"][0][k] - ref).abs().max())
        assert err <= 2e-2 * max(1.0, float(ref.abs().max())), (k, err)
    num = den = 0.0
    for k, gref in res["fp32"][1].items():
        gr = res["bf16"][1][k]
        assert gr.shape == gref.shape and torch.isfinite(gr).all(), k
        r = float((gr - gref).norm() / (gref.norm() + 1e-30))
        num += float((gr - gref).double().pow(2).sum())
        den += float(gref.double().pow(2).sum())
        assert r <= 0.1, (k, r)        # first layer: ReLU sign flips of bf16-rounded pre-activations accumulate (cf. test_baseline_size_gpu.py)
    whole = (num / den) ** 0.5
    print(f"[tnet bf16 R={R}] whole-network gradient error vs torch fp32: {whole:.2e}")
    assert whole <= 4e-2, whole


def test_tnet_partial_cotangents_and_sink(cuda_dev):
    """The training step reads only t_beta and t_alpha (losses.py:53-60): rgb gets no gradient, rgb_layer's
    parameters stay at zero gradient; with `grad_sink` the kernels accumulate into existing .grad buffers."""
    R, n_img = 512, 9
    b = synth.ray_batch(R, n_img, 72)
    feats, idx = b["feats"].to(cuda_dev), b["img_idx"].to(cuda_dev)
    ref = _net(n_img, 3, cuda_dev, "fp32")
    o = ref(feats, idx)
    (o["beta"].sum() * 0.5 + (o["alpha"] ** 2).sum()).backward()
    net = _net(n_img, 3, cuda_dev, "bf16")
    net.grad_sink = True
    for p in net.parameters():
        p.grad = torch.ones_like(p)                 # pre-existing content must be ADDED to
    o = net(feats, idx)
    (o["beta"].sum() * 0.5 + (o["alpha"] ** 2).sum()).backward()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        got = p.grad - 1.0
        if k.startswith("rgb_layer"):
            assert q.grad is None and float(got.abs().max()) == 0.0, k
            continue
        r = float((got - q.grad).norm() / (q.grad.norm() + 1e-30))
        assert r <= 0.1, (k, r)
