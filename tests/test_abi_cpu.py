"""CPU: the C-ABI shared library builds, loads and exports every symbol include/*.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def libpath():
    from upnerf_b200 import build

    return build.build()


def test_exports_match_header(libpath):
    hdr = (ROOT / "include" / "upnerf_b200.h").read_text()
    declared = set(re.findall(r"\b(upnerf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(str(libpath))
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header(libpath, tmp_path):
    """ctypes mirrors must have the sizes the C compiler gives the header structs."""
    import subprocess

    from upnerf_b200 import _lib as L

    src = tmp_path / "sz.cpp"
    src.write_text('#include "upnerf_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(upnerf_render_args),sizeof(upnerf_composite_args),sizeof(upnerf_net_config),"
                   "sizeof(upnerf_pass_io),sizeof(upnerf_epilogue),sizeof(upnerf_ray_batch_args),"
                   "sizeof(upnerf_tail_args),sizeof(upnerf_adam_args));}\n")
    exe = tmp_path / "sz"
    subprocess.run(["g++", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(t) for t in (L.RenderArgs, L.CompositeArgs, L.NetConfig, L.PassIO, L.Epilogue, L.RayBatchArgs,
                                     L.TailArgs, L.AdamArgs)]
    assert got == want


def test_param_count_and_errors(libpath):
    from upnerf_b200 import _lib as L

    cfg = L.NetConfig(8, 256, 10, 4, 1, 384, 48, 16, 1, 1, 1, 0.1, 0.5)
    assert L.nerf_param_count(cfg) == 818182          # SURVEY.md 8(a6), probed on the reference
    cfg1 = L.NetConfig(8, 256, 10, 4, 0, 0, 0, 0, 0, 0, 0, 0.0, 1.0)
    assert L.nerf_param_count(cfg1) == 595845         # config 1 static MLP
    bad = L.NetConfig(8, 128, 10, 4, 1, 384, 48, 16, 1, 1, 1, 0.1, 0.5)
    with pytest.raises(L.UpnerfError):
        L.nerf_param_count(bad)


def test_module_state_dict_matches_reference_layout():
    import torch

    from oracle import synth
    from oracle.upnerf_oracle import NerfConfig
    from upnerf_b200.models.nerf import NeRF, flat_parameters

    for kw in (dict(), dict(encode_feat=False, feat_dim=0, appearance_dim=0, candidate_dim=0)):
        cfg = NerfConfig(**kw)
        m = NeRF("coarse", encode_feat=cfg.encode_feat, feat_dim=cfg.feat_dim, xyz_L=10, dir_L=4,
                 appearance_dim=cfg.appearance_dim, candidate_dim=cfg.candidate_dim, c2f=(0.1, 0.5))
        want = synth.nerf_param_shapes(cfg)
        got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert got == want
        assert flat_parameters(m).numel() == sum(int(torch.tensor(s).prod()) if s else 1 for _, s in want)


def test_cpu_tensors_are_rejected():
    import torch

    from upnerf_b200 import _lib as L
    from upnerf_b200.models.rendering import render_rays

    with pytest.raises(L.UpnerfError):
        render_rays({}, {}, torch.zeros(4, 8), torch.zeros(4, dtype=torch.long), 1.0)


def test_adam_class_table():
    from upnerf_b200.models.nerf_system import adam_class

    assert adam_class("nerf_coarse.progress") == "never"
    assert adam_class("nerf_fine.xyz_encoding_3.0.weight") == "always"
    assert adam_class("nerf_fine.feat_share_layer.bias") == "always"
    assert adam_class("nerf_coarse.rgb_share_layer.2.weight") == "rgb"
    assert adam_class("nerf_coarse.candidate_encoding.0.weight") == "cand"
    assert adam_class("nerf_coarse.candidate_sigma.0.bias") == "cand"
    assert adam_class("nerf_coarse.feat_candidate_layer.weight") == "cand"
    assert adam_class("embedding_fine_a.weight") == "rgb"
    assert adam_class("embedding_coarse_c.weight") == "cand"
    assert adam_class("transient_net.rgb_layer.0.weight") == "never"
    assert adam_class("transient_net.feat_encoder.0.weight") == "rgb"
    assert adam_class("depth_scale.weight") == "cand"
    assert adam_class("se3_refine.weight") == "always"


def test_host_helpers_reject_cpu():
    """The resident batcher and the host pipeline are CUDA-only like the rest of the path: they raise
    instead of silently running somewhere else."""
    import pytest
    import torch

    from oracle import ray_batch as RB
    from upnerf_b200 import _lib as L
    from upnerf_b200.datasets import RayBatcher
    from upnerf_b200.utils.pipeline import DevicePrefetcher

    t = RB.synth_tables(2, 4, 4, 3, 8, seed=1)
    with pytest.raises(L.UpnerfError):
        RayBatcher(t["all_ray_infos"], t["all_directions"], t["all_rgbs"], t["poses"], device="cpu")
    with pytest.raises(RuntimeError):
        DevicePrefetcher([], "cpu")


def test_ray_batch_oracle_edge_cases():
    """Oracle restatement of the sample path: empty index set, repeated indices, the four corners."""
    import torch

    from oracle import ray_batch as RB

    t = RB.synth_tables(2, 5, 7, 4, 6, seed=3)
    empty = RB.getitem_batch(t, torch.zeros(0, dtype=torch.int64))
    assert empty["feats"].shape == (0, 6) and empty["c2w"].shape == (0, 3, 4)
    n = 5 * 7
    corners = torch.tensor([0, 6, n - 7, n - 1, n, 2 * n - 1, 3, 3])
    out = RB.getitem_batch(t, corners)
    assert torch.equal(out["img_idx"], torch.tensor([0, 0, 0, 0, 1, 1, 0, 0]))
    assert torch.equal(out["feats"][6], out["feats"][7])
    # top-left corner = the map's (0, 0) pixel exactly; any sample on the last row / column = 0
    assert torch.equal(out["feats"][0], t["feat_maps"][0, 0, 0])
    assert (out["feats"][[1, 2, 3, 5]] == 0).all()
