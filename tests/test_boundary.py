"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the host mirror keeps the reference's names / signatures / state_dict keys, and the product
package never touches the oracle or falls back to CPU."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import fv2p_b200
from fv2p_b200 import _lib, spconv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "from-voxel-to-point_b200")


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "fv2p_b200.h")).read()
    return sorted(set(re.findall(r"FV2P_API[^;(]*?\b(fv2p_\w+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    syms = _header_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.PROTOTYPES) == syms  # the ctypes table binds exactly the header
    assert lib.fv2p_abi_version() == 2


def test_size_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.fv2p_voxelize_workspace_bytes(100000, 4, 30000, 5, 100000) > 100000 * 4
    assert lib.fv2p_rulebook_workspace_bytes(1000, 8000, 27) > 27 * 1000 * 4
    assert lib.fv2p_rulebook_workspace_bytes(1000, 8000, 64) == 0  # kernel volume above FV2P_MAX_KVOL
    assert lib.fv2p_indice_conv_workspace_bytes(27, 1000) >= 27 * 1000 * 4


def test_invalid_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    three = _lib.i32x3([3, 3, 3])
    st = lib.fv2p_rulebook_subm(None, 10, None, 1, three, _lib.i32x3([9, 9, 9]), three, None, 0, None, None, 0,
                                None, 0, None, None)
    assert st == -1 and b"kernel volume" in lib.fv2p_last_error()
    st = lib.fv2p_conv_fwd(None, 0, None, None, 0, None, None, None, 99, 0, None, 4, 4, None, None, None, None, 0, 0, None, None)
    assert st == -1
    with pytest.raises(ValueError):
        _lib.check(st, "conv_fwd")


def test_cpu_tensors_are_refused_not_emulated():
    ind = torch.zeros((4, 4), dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA"):
        spconv.ops.get_indice_pairs(ind, 1, [8, 8, 8], 3, 1, 1, 1, 0, True)
    x = spconv.SparseConvTensor(torch.zeros(4, 4), ind, [8, 8, 8], 1)
    with pytest.raises(ValueError, match="CUDA"):
        spconv.SubMConv3d(4, 8, 3, indice_key="k")(x)
    with pytest.raises(ValueError, match="CUDA"):
        fv2p_b200.MeanVFE({}, 4)({"voxels": torch.zeros(2, 5, 4), "voxel_num_points": torch.ones(2)})


def _mentions_oracle_code(text):
    return bool(re.search(r"^\s*(from|import)\s+\.*oracle|oracle[/\\.](oracle|ref|_ref|_build|build_ref)|"
                          r"libfv2p_oracle|orc_\w+\(", text, flags=re.M))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not _mentions_oracle_code(text), os.path.join(dirpath, f)
    assert not _mentions_oracle_code(open(os.path.join(ROOT, "fv2p_b200.py")).read())


def test_module_api_matches_reference_signatures():
    sig = inspect.signature(spconv.SubMConv3d.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation",
                                        "groups", "bias", "indice_key"]
    assert list(inspect.signature(spconv.SparseConv3d.__init__).parameters) == list(sig.parameters)
    assert list(inspect.signature(spconv.SparseConvTensor.__init__).parameters)[1:] == [
        "features", "indices", "spatial_shape", "batch_size", "grid"]
    assert list(inspect.signature(fv2p_b200.VoxelGenerator.__init__).parameters)[1:5] == [
        "voxel_size", "point_cloud_range", "max_num_points", "max_voxels"]
    conv = spconv.SparseConv3d(4, 8, (3, 1, 1), stride=(2, 1, 1), bias=True)
    assert tuple(conv.weight.shape) == (3, 1, 1, 4, 8) and tuple(conv.bias.shape) == (8,)
    seq = spconv.SparseSequential(conv, torch.nn.BatchNorm1d(8), torch.nn.ReLU())
    assert list(seq.state_dict())[:3] == ["0.weight", "0.bias", "1.weight"]
    for name in ("SparseInverseConv3d", "SparseMaxPool3d", "SubMConv2d"):
        with pytest.raises(NotImplementedError):
            getattr(spconv, name)(4, 4, 3)


@pytest.mark.parametrize("name,nkeys,c4", [("VoxelBackBone8x", 72, 64), ("VoxelResBackBone8x", 142, 128)])
def test_backbone_state_dict_keys_and_shapes(name, nkeys, c4):
    net = getattr(fv2p_b200, name)({}, 4, np.array([1408, 1600, 40]))
    sd = net.state_dict()
    assert len(sd) == nkeys
    assert tuple(sd["conv_input.0.weight"].shape) == (3, 3, 3, 4, 16)
    assert tuple(sd["conv_out.0.weight"].shape) == (3, 1, 1, c4, 128)
    assert "conv_input.1.running_var" in sd and "conv_input.1.num_batches_tracked" in sd
    assert net.sparse_shape.tolist() == [41, 1600, 1408]
    if name == "VoxelResBackBone8x":
        assert tuple(sd["conv4.1.conv1.bias"].shape) == (128,) and net.num_point_features == 128
    else:
        assert net.num_point_features == {"x_conv1": 16, "x_conv2": 32, "x_conv3": 64, "x_conv4": 64}
    assert fv2p_b200.BACKBONES_3D[name] is type(net)


def test_engine_trace_matches_layer_plan():
    from fv2p_b200.engine import trace_backbone
    steps, books, levels = trace_backbone(fv2p_b200.VoxelResBackBone8x({}, 5, np.array([1504, 1504, 40])))
    assert len(steps) == 21 and [b.key for b in books] == ["subm1", "res1", "spconv2", "res2", "spconv3", "res3",
                                                           "spconv4", "res4", "spconv_down2"]
    assert levels == [[41, 1504, 1504], [21, 752, 752], [11, 376, 376], [5, 188, 188], [2, 188, 188]]
    assert [s.export for s in steps if s.export] == ["x_conv1", "x_conv2", "x_conv3", "x_conv4", "out"]
    assert sum(s.res_buf is not None for s in steps) == 8


def test_candidate_fanout():
    f = spconv.ops.candidate_fanout
    assert f([3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1]) == 8
    assert f([3, 1, 1], [2, 1, 1], [0, 0, 0], [1, 1, 1]) == 2
    assert f([3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1]) == 27
    assert f([2, 2, 2], [2, 2, 2], [0, 0, 0], [1, 1, 1]) == 1


def test_arena_capacity_bounds_grow_towards_the_hard_bound():
    """Host logic of the row-capacity bounds (engine._caps / grow): modest bounds first, x4 per overflow, then the
    hard bound min(fan-out * N, dense volume); never above the hard bound."""
    import numpy as np
    import fv2p_b200
    from fv2p_b200 import synth
    cfg = synth.DATASETS["kitti"]
    net = fv2p_b200.VoxelResBackBone8x({"CAP_GROWTH": 2.0}, 4, np.array(synth.grid_size(cfg))).eval()
    eng = net.get_engine()
    hard = fv2p_b200.BackboneEngine(net, cap_growth=None)._caps(100000, 8)
    assert hard == [100000, 800000, 6400000, 1408000, 563200]
    first = eng._caps(100000, 8)
    assert first[0] == 100000 and all(a <= b for a, b in zip(first, hard))
    assert first[1] == 2 * 100000 + 1024 and first[2] == 2 * first[1] + 1024
    seen = [eng.cap_growth]
    while eng.grow():
        seen.append(eng.cap_growth)
        assert all(a <= b for a, b in zip(eng._caps(100000, 8), hard))
    assert seen == [2.0, 8.0, None] and eng._caps(100000, 8) == hard and eng.arena is None
