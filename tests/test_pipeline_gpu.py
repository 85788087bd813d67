"""Host pipeline helpers (upnerf_b200/utils/pipeline.py) on the GPU: the prefetcher hands out the batches it was
given, in order, from two alternating preallocated device sets; the delayed scalar returns every pushed value
exactly once, in order, `depth` calls late."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_prefetcher_order_and_reuse(cuda_dev):
    from upnerf_b200.utils.pipeline import DevicePrefetcher

    n = 7
    host = [{"a": torch.full((1000, 3), float(i)).pin_memory(), "i": torch.full((1000,), i, dtype=torch.int64).pin_memory()}
            for i in range(n)]
    ptrs, seen = set(), []
    acc = torch.zeros((), device=cuda_dev)
    for j, b in enumerate(DevicePrefetcher(iter(host), cuda_dev)):
        assert b["a"].device.type == "cuda" and b["i"].dtype == torch.int64
        ptrs.add(b["a"].data_ptr())
        # consume the batch with queued device work only (no host sync inside the loop): the next-but-one batch
        # overwrites this buffer and must wait for this kernel on the device
        acc = acc + b["a"].sum() / 3000.0 + (b["i"].float().mean() - j).abs()
        seen.append(b["a"][0, 0])
    torch.cuda.synchronize()
    assert len(seen) == n and len(ptrs) == 2                 # two device sets, alternating
    assert float(acc) == sum(range(n))                       # every batch arrived intact and in order


def test_device_prefetcher_rejects_cpu():
    from upnerf_b200.utils.pipeline import DevicePrefetcher

    with pytest.raises(RuntimeError):
        DevicePrefetcher(iter([]), "cpu")


@pytest.mark.parametrize("depth", [1, 3])
def test_delayed_scalar_returns_every_value_once(cuda_dev, depth):
    from upnerf_b200.utils.pipeline import DelayedScalar

    d = DelayedScalar(depth=depth)
    got = []
    for i in range(10):
        v = d.push(torch.tensor(float(i), device=cuda_dev) * 2)
        assert (v is None) == (i < depth)
        if v is not None:
            got.append(v)
    got.extend(d.drain())
    assert got == [2.0 * i for i in range(10)]
    assert d.drain() == [] and d.last() is None
