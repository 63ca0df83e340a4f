"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed reference
fixtures.  Bar: voxel coordinates, counts and rulebooks BIT-EXACT; features within 1e-4 relative
(max|a-b|/max|b| per tensor) on the fp32 path, within 3e-2 on the bf16 path."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import oracle as O

import fv2p_b200
from fv2p_b200 import _lib, spconv, synth

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 3e-2
DEV = "cuda:0"


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.to(dtype) if dtype is not None else t


def test_native_library_is_loaded_and_device_is_sm100():
    lib = _lib.load()
    sm, major, minor = (_lib.ctypes.c_int(0) for _ in range(3))
    st = lib.fv2p_device_check(_lib.ctypes.byref(sm), _lib.ctypes.byref(major), _lib.ctypes.byref(minor))
    assert st == 0, _lib.last_error()
    assert major.value == 10 and sm.value >= 100


# ------------------------------------------------------------------------------------- voxelizer
def test_voxelizer_matches_reference_fixtures():
    g = load_golden("voxelize")
    for i in range(int(g["n_cases"])):
        vg = fv2p_b200.VoxelGenerator(g[f"c{i}_voxel_size"], g[f"c{i}_range"], int(g[f"c{i}_T"]),
                                      int(g[f"c{i}_max_voxels"]))
        voxels, coors, num = vg.generate(g[f"c{i}_points"])
        assert np.array_equal(coors, g[f"c{i}_coors"]), i
        assert np.array_equal(num, g[f"c{i}_num"]), i
        assert np.array_equal(voxels, g[f"c{i}_voxels"]), i
        assert rel_err(vg.last_voxel_features.cpu().numpy(), g[f"c{i}_mean"]) < 1e-6, i
        # MeanVFE module on the padded tensor (reference batch_dict contract: counts arrive as floats)
        bd = fv2p_b200.MeanVFE({}, voxels.shape[2])({"voxels": cuda(voxels), "voxel_num_points": cuda(num).float()})
        assert rel_err(bd["voxel_features"].cpu().numpy(), g[f"c{i}_mean"]) < 1e-6, i


@pytest.mark.parametrize("ds,shuffle,max_voxels", [("kitti", False, 40000), ("kitti", True, 16000),
                                                   ("kitti", True, 5000), ("waymo", False, 90000),
                                                   ("waymo", True, 30000)])
def test_voxelizer_full_frames_match_oracle(ds, shuffle, max_voxels):
    cfg = synth.DATASETS[ds]
    pts = synth.lidar_frame(ds, seed=11, shuffle=shuffle)
    ref = O.voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels)
    vg = fv2p_b200.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels)
    got = vg.generate(pts)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert rel_err(vg.last_voxel_features.cpu().numpy(), O.mean_vfe(ref[0], ref[2])) < 1e-6


def test_batch_voxelizer_ragged_batch_and_empty_frame():
    cfg = synth.DATASETS["kitti"]
    frames = [synth.lidar_frame("kitti", seed=s, az_steps=a, shuffle=sh)
              for s, a, sh in ((1, 300, False), (2, 40, True), (3, 120, False))]
    frames.insert(2, np.zeros((0, 4), np.float32))  # an empty frame in the middle
    max_voxels = 9000
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int32)
    bv = fv2p_b200.BatchVoxelizer(cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels, want_voxels=True)
    out = bv(cuda(np.concatenate(frames)), cuda(offs), max(f.shape[0] for f in frames))
    voff = out["voxel_offsets"].cpu().numpy()
    assert int(out["status"].item()) == 0
    exp_coords, exp_feat, exp_num = [], [], []
    for b, f in enumerate(frames):
        v, c, n = O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels)
        assert voff[b + 1] - voff[b] == c.shape[0]
        exp_coords.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
        exp_feat.append(O.mean_vfe(v, n))
        exp_num.append(n)
    m = voff[-1]
    assert np.array_equal(out["voxel_coords"][:m].cpu().numpy(), np.concatenate(exp_coords))
    assert np.array_equal(out["voxel_num_points"][:m].cpu().numpy(), np.concatenate(exp_num))
    assert rel_err(out["voxel_features"][:m].cpu().numpy(), np.concatenate(exp_feat)) < 1e-6


def test_voxelizer_is_deterministic_and_permutation_sensitive_like_the_reference():
    cfg = synth.DATASETS["waymo"]
    pts = synth.lidar_frame("waymo", seed=5, az_steps=900)
    vg = fv2p_b200.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], 5, 20000)
    a = vg.generate(pts)
    b = vg.generate(pts)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # the set of voxels below the cap is order dependent (first arrival), the oracle agrees on a permutation
    perm = np.random.default_rng(0).permutation(pts.shape[0])
    c = vg.generate(pts[perm])
    r = O.voxelize(pts[perm], cfg["voxel_size"], cfg["point_cloud_range"], 5, 20000)
    for x, y in zip(c, r):
        assert np.array_equal(x, y)


# ------------------------------------------------------------------------------------- rulebooks
GEOMS = {"subm3": (True, 3, 1, 1), "s2p1": (False, 3, 2, 1), "s2p011": (False, 3, 2, (0, 1, 1)),
         "down311": (False, (3, 1, 1), (2, 1, 1), 0), "s1p1": (False, 3, 1, 1), "k2s2": (False, 2, 2, 0)}


@pytest.mark.parametrize("cset", ["rand", "shuf", "blob"])
@pytest.mark.parametrize("geom", sorted(GEOMS))
def test_rulebooks_match_reference_fixtures(cset, geom):
    g = load_golden("rulebook_conv")
    subm, ks, st, pd = GEOMS[geom]
    ind = cuda(g[f"{cset}_indices"])
    outids, pairs, num, nbr = spconv.ops.get_indice_pairs(ind, int(g[f"{cset}_batch"]), g["shape"].tolist(), ks, st,
                                                          pd, 1, 0, subm, False, return_nbr=True)
    assert np.array_equal(outids.cpu().numpy(), g[f"{cset}_{geom}_outids"])
    assert np.array_equal(num.cpu().numpy(), g[f"{cset}_{geom}_num"])
    assert np.array_equal(pairs.cpu().numpy(), g[f"{cset}_{geom}_pairs"])
    # the neighbour map carries the same pairs, output-major
    p, n, m = pairs.cpu().numpy(), num.cpu().numpy(), nbr.cpu().numpy()
    assert (m >= 0).sum() == n.sum()
    for k in range(p.shape[0]):
        assert np.array_equal(m[k][p[k, 1, :n[k]]], p[k, 0, :n[k]])
    assert np.array_equal(spconv.ops.pairs_to_nbr(pairs, num, outids.shape[0]).cpu().numpy(), m)


@pytest.mark.parametrize("ds", ["kitti", "waymo"])
def test_rulebooks_full_frame_chain_matches_oracle(ds):
    """Every rulebook of the backbone, level by level, on a full-size frame (bit-exact incl. the -1 tail)."""
    cfg = synth.DATASETS[ds]
    pts = synth.lidar_frame(ds, seed=21)
    _, c, _ = O.voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], 5, cfg["max_voxels"]["test"])
    ind = O.collate([c])
    gs = synth.grid_size(cfg)
    shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
    chain = [(True, 3, 1, 1), (False, 3, 2, 1), (True, 3, 1, 1), (False, 3, 2, 1), (True, 3, 1, 1),
             (False, 3, 2, (0, 1, 1)), (True, 3, 1, 1), (False, (3, 1, 1), (2, 1, 1), 0)]
    for subm, ks, st, pd in chain:
        got = spconv.ops.get_indice_pairs(cuda(ind), 1, shape, ks, st, pd, 1, 0, subm, False)
        if subm:
            ref = O.rulebook_subm(ind, 1, shape, ks, 1)
        else:
            ref = O.rulebook_conv(ind, 1, shape, ks, st, pd, 1)
        for a, b in zip(got, ref[:3]):
            assert np.array_equal(a.cpu().numpy(), b)
        if not subm:
            ind, shape = ref[0], ref[3]


def test_rulebook_batch_above_int32_grid_limit():
    """64 frames in one call: the reference's dense grid overflows int32 past 23 frames;
    the hash path must equal the per-frame oracle with batch offsets."""
    shape = [41, 1600, 1408]
    base = synth.random_voxels(shape, 300, 1, seed=3)
    frames = []
    for b in range(64):
        f = base.copy()
        f[:, 0] = b
        frames.append(f)
    ind = np.concatenate(frames)
    _, pairs, num = spconv.ops.get_indice_pairs(cuda(ind), 64, shape, 3, 1, 1, 1, 0, True, False)
    _, p1, n1 = O.rulebook_subm(base, 1, shape, 3, 1)
    assert np.array_equal(num.cpu().numpy(), n1 * 64)
    outids, pairs, num = spconv.ops.get_indice_pairs(cuda(ind), 64, shape, 3, 2, 1, 1, 0, False, False)
    o1, _, n1, _ = O.rulebook_conv(base, 1, shape, 3, 2, 1, 1)
    assert np.array_equal(num.cpu().numpy(), n1 * 64)
    assert np.array_equal(outids.cpu().numpy()[-o1.shape[0]:, 1:], o1[:, 1:])
    assert outids.shape[0] == 64 * o1.shape[0]


def test_rulebook_empty_and_single_voxel():
    empty = torch.zeros((0, 4), dtype=torch.int32, device=DEV)
    outids, pairs, num = spconv.ops.get_indice_pairs(empty, 1, [8, 8, 8], 3, 1, 1, 1, 0, True, False)
    assert pairs.shape == (27, 2, 0) and int(num.sum()) == 0
    outids, pairs, num = spconv.ops.get_indice_pairs(empty, 1, [8, 8, 8], 3, 2, 1, 1, 0, False, False)
    assert outids.shape[0] == 0 and int(num.sum()) == 0
    one = np.int32([[0, 0, 0, 0]])
    got = spconv.ops.get_indice_pairs(cuda(one), 1, [8, 8, 8], 3, 2, 1, 1, 0, False, False)
    ref = O.rulebook_conv(one, 1, [8, 8, 8], 3, 2, 1, 1)
    for a, b in zip(got, ref[:3]):
        assert np.array_equal(a.cpu().numpy(), b)


def test_rulebook_dilated_and_even_submanifold_geometries():
    """Non mirror-symmetric submanifold kernels take the input-side probe path."""
    ind = synth.random_voxels([9, 20, 20], 900, 2, seed=8)
    for ks, dil in ((3, 2), ((2, 2, 2), 1), ((1, 3, 3), 1)):
        got = spconv.ops.get_indice_pairs(cuda(ind), 2, [9, 20, 20], ks, 1, 0, dil, 0, True, False)
        ref = O.rulebook_subm(ind, 2, [9, 20, 20], ks, dil)
        for a, b in zip(got, ref):
            assert np.array_equal(a.cpu().numpy(), b), (ks, dil)


# ------------------------------------------------------------------------------------- convolution
@pytest.mark.parametrize("geom", sorted(GEOMS))
@pytest.mark.parametrize("ch", [(4, 16), (16, 32), (5, 16)])
def test_indice_conv_matches_reference_fixtures(geom, ch):
    g = load_golden("rulebook_conv")
    key = f"conv_{geom}_{ch[0]}_{ch[1]}"
    out = spconv.ops.indice_conv(cuda(g[key + "_feats"]), cuda(g[key + "_w"]), cuda(g[f"blob_{geom}_pairs"]),
                                 cuda(g[f"blob_{geom}_num"]), g[f"blob_{geom}_outids"].shape[0], False, GEOMS[geom][0])
    assert rel_err(out.cpu().numpy(), g[key + "_out"]) < FP32_TOL


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 64), (64, 64), (128, 128), (7, 19), (64, 128)])
def test_conv_fused_epilogue_matches_oracle(cin, cout):
    rng = np.random.default_rng(cin * 1000 + cout)
    ind = synth.random_voxels([7, 24, 24], 1500, 2, seed=cin)
    feats = rng.standard_normal((ind.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    bias, gamma, beta, mean = (rng.standard_normal(cout).astype(np.float32) * 0.1 for _ in range(4))
    var = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    res = rng.standard_normal((ind.shape[0], cout)).astype(np.float32)
    _, pairs, num = O.rulebook_subm(ind, 2, [7, 24, 24], 3, 1)
    ref = O.bias_bn_res_relu(O.indice_conv(feats, w, pairs, num, ind.shape[0], False, True), bias,
                             (gamma + 1, beta, mean, var), res, True)
    _, _, _, nbr = spconv.ops.get_indice_pairs(cuda(ind), 2, [7, 24, 24], 3, 1, 1, 1, 0, True, False, return_nbr=True)
    scale = (gamma + 1) / np.sqrt(var + 1e-3)
    shift = beta - mean * scale
    out = spconv.ops.conv_forward(cuda(feats), cuda(w), nbr, ind.shape[0], cuda(bias), cuda(scale.astype(np.float32)),
                                  cuda(shift.astype(np.float32)), cuda(res), True)
    assert rel_err(out.cpu().numpy(), ref) < FP32_TOL


@pytest.mark.parametrize("mode", ["tf32x3", "bf16"])
@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)])
@pytest.mark.parametrize("geom", ["subm", "s2"])
def test_tensor_core_conv_matches_oracle(mode, cin, cout, geom):
    """tcgen05 kernels: the fp32 mode (split-bf16 products) must hold the fp32 bar, bf16 its stated tolerance, fused
    epilogue included."""
    rng = np.random.default_rng(cin * 7 + cout)
    shape = [9, 40, 40]
    ind = synth.random_voxels(shape, 2200, 2, seed=cout)
    # clustered coordinates so that tiles have many active offsets and missing neighbours at the same time
    ind[:, 1:] = ind[:, 1:] // np.array([2, 3, 3])
    ind = np.unique(ind, axis=0).astype(np.int32)
    feats = rng.standard_normal((ind.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    bias, beta, mean = (rng.standard_normal(cout).astype(np.float32) * 0.1 for _ in range(3))
    gamma = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    var = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    if geom == "subm":
        outids, pairs, num = O.rulebook_subm(ind, 2, shape, 3, 1)
        got_rb = spconv.ops.get_indice_pairs(cuda(ind), 2, shape, 3, 1, 1, 1, 0, True, False, return_nbr=True)
    else:
        outids, pairs, num, _ = O.rulebook_conv(ind, 2, shape, 3, 2, 1, 1)
        got_rb = spconv.ops.get_indice_pairs(cuda(ind), 2, shape, 3, 2, 1, 1, 0, False, False, return_nbr=True)
    n_out = outids.shape[0]
    res = rng.standard_normal((n_out, cout)).astype(np.float32)
    nbr = got_rb[3].contiguous()
    scale = (gamma / np.sqrt(var + 1e-3)).astype(np.float32)
    shift = (beta - mean * scale).astype(np.float32)
    if mode == "bf16":
        q = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()  # noqa: E731
        ref = O.bias_bn_res_relu(O.indice_conv(q(feats), q(w), pairs, num, n_out, False, geom == "subm"), bias,
                                 (gamma, beta, mean, var), q(res), True)
        packed = spconv.ops.pack_weight(cuda(w), _lib.MODE_BF16_TC)
        out = spconv.ops.conv_forward(cuda(feats, torch.bfloat16), packed, nbr, n_out, cuda(bias), cuda(scale),
                                      cuda(shift), cuda(res, torch.bfloat16), True, mode=_lib.MODE_BF16_TC)
        assert out.dtype == torch.bfloat16
        assert rel_err(out.float().cpu().numpy(), ref) < 1e-2  # one bf16 rounding of the output
    else:
        # float64 contraction as the truth here: the fp32 oracle's own summation error over 27*cin terms is of
        # the same order as the bar being checked
        acc = np.zeros((n_out, cout), np.float64)
        for k in range(27):
            t = num[k]
            np.add.at(acc, pairs[k, 1, :t], feats[pairs[k, 0, :t]].astype(np.float64) @ w[k].astype(np.float64))
        assert rel_err(O.indice_conv(feats, w, pairs, num, n_out, False, geom == "subm"), acc) < 1e-5
        ref = O.bias_bn_res_relu(acc.astype(np.float32), bias, (gamma, beta, mean, var), res, True)
        packed = spconv.ops.pack_weight(cuda(w), _lib.MODE_TF32X3_TC)
        out = spconv.ops.conv_forward(cuda(feats), packed, nbr, n_out, cuda(bias), cuda(scale), cuda(shift),
                                      cuda(res), True, mode=_lib.MODE_TF32X3_TC)
        # measured 3.5e-6 ... 1.3e-5 at 27*16 ... 27*128 terms (the tensor core adds each MMA's products into the fp32
        # accumulator with round-toward-zero, a bias that grows with the number of accumulating MMAs): 8x inside 1e-4
        err = rel_err(out.cpu().numpy(), ref)
        print("fp32 tensor-core conv vs fp64 truth: %s %d->%d rel err %.2e" % (geom, cin, cout, err))
        assert err < 5e-5, err


def test_sparse_conv_equals_dense_conv3d():
    """The property the reference's test_utils.py:144-193 was written for: sparse conv == F.conv3d on .dense()."""
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False  # the dense comparison must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    shape = [6, 10, 9]
    ind = synth.random_voxels(shape, 200, 2, seed=4)
    feats = torch.randn(ind.shape[0], 8, device=DEV)
    x = spconv.SparseConvTensor(feats, cuda(ind), shape, 2)
    for conv in (spconv.SparseConv3d(8, 12, 3, stride=2, padding=1, bias=True),
                 spconv.SparseConv3d(8, 12, (3, 1, 1), stride=(2, 1, 1), padding=0, bias=False),
                 spconv.SparseConv3d(8, 12, 3, stride=1, padding=1, bias=True)):
        conv = conv.to(DEV)
        y = conv(x)
        w = conv.weight.permute(4, 3, 0, 1, 2).contiguous()
        ref = torch.nn.functional.conv3d(x.dense(), w, conv.bias, conv.stride, conv.padding)
        dense = y.dense()
        # a strided sparse conv only materialises outputs that have at least one active input
        mask = (y.dense().abs().sum(1, keepdim=True) > 0) | (ref.abs().sum(1, keepdim=True) == 0)
        assert dense.shape == ref.shape
        assert torch.allclose(dense * mask, ref * mask, atol=1e-4, rtol=1e-4)
    sub = spconv.SubMConv3d(8, 12, 3, bias=False, indice_key="s").to(DEV)
    y = sub(x)
    w = sub.weight.permute(4, 3, 0, 1, 2).contiguous()
    ref = torch.nn.functional.conv3d(x.dense(), w, None, 1, 1)
    active = (x.dense().abs().sum(1, keepdim=True) > 0).float()
    assert torch.allclose(y.dense(), ref * active, atol=1e-4, rtol=1e-4)


# ------------------------------------------------------------------------------------- backbones
def _load_backbone(name, in_ch, grid, seed, cfg=None):
    net = getattr(fv2p_b200, name)(dict(cfg or {}), in_ch, np.array(grid)).eval()
    state = synth.randomize_state(net.state_dict(), seed=seed)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=False)
    return net.to(DEV), state


@pytest.mark.parametrize("ds", ["kitti", "waymo"])
@pytest.mark.parametrize("name", ["VoxelBackBone8x", "VoxelResBackBone8x"])
@pytest.mark.parametrize("fused", [True, False])
def test_backbone_matches_reference_fixtures(ds, name, fused):
    g = load_golden(f"backbone_{ds}_{name}")
    net, _ = _load_backbone(name, g["voxel_features"].shape[1], g["grid_size"], int(g["seed"]), {"FUSED": fused})
    bd = {"voxel_features": cuda(g["voxel_features"]), "voxel_coords": cuda(g["voxel_coords"]).float(),
          "batch_size": int(g["batch_size"])}
    with torch.no_grad():
        bd = net(bd)
    outs = dict(bd["multi_scale_3d_features"], out=bd["encoded_spconv_tensor"])
    for k in ("x_conv1", "x_conv2", "x_conv3", "x_conv4", "out"):
        assert np.array_equal(outs[k].indices.cpu().numpy(), g[k + "_indices"]), k
        assert rel_err(outs[k].features.float().cpu().numpy(), g[k + "_features"]) < FP32_TOL, k
    assert list(outs["out"].spatial_shape) == g["out_shape"].tolist()
    idict = outs["out"].indice_dict
    for key in [f[3:-4] for f in g.files if f.startswith("rb_") and f.endswith("_num")]:
        outids, _, pairs, num, _ = idict[key]
        assert np.array_equal(num.cpu().numpy(), g[f"rb_{key}_num"]), key
        assert outids.shape[0] == int(g[f"rb_{key}_nout"]), key
        assert pairs.shape[0] == num.shape[0] and pairs.shape[1] == 2


@pytest.mark.parametrize("name", ["VoxelBackBone8x", "VoxelResBackBone8x"])
def test_backbone_full_kitti_frame_vs_oracle_and_rulebooks_bit_exact(name):
    cfg = synth.DATASETS["kitti"]
    pts = synth.lidar_frame("kitti", seed=31)
    v, c, n = O.voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], 5, 40000)
    feats, coords = O.mean_vfe(v, n), O.collate([c])
    gs = synth.grid_size(cfg)
    net, state = _load_backbone(name, 4, gs, 5)
    ref = O.backbone_forward(name, state, feats, coords, 1, [int(gs[2]) + 1, int(gs[1]), int(gs[0])])
    with torch.no_grad():
        bd = net({"voxel_features": cuda(feats), "voxel_coords": cuda(coords), "batch_size": 1})
    outs = dict(bd["multi_scale_3d_features"], out=bd["encoded_spconv_tensor"])
    for k in ("x_conv1", "x_conv2", "x_conv3", "x_conv4", "out"):
        assert np.array_equal(outs[k].indices.cpu().numpy(), ref[k][1]), k
        assert rel_err(outs[k].features.cpu().numpy(), ref[k][0]) < FP32_TOL, k
    for key, (outids, pairs, num) in ref["rulebooks"].items():
        g_out, _, g_pairs, g_num, _ = outs["out"].indice_dict[key]
        assert np.array_equal(g_out.cpu().numpy(), outids), key
        assert np.array_equal(g_num.cpu().numpy(), num), key
        assert np.array_equal(g_pairs.cpu().numpy(), pairs), key


def test_hot_path_from_host_points_batch_invariance():
    """Whole path from host buffers: frames processed in a batch equal the same frames processed alone
    (frames are independent units -- the multi-GPU sharding relies on exactly this)."""
    cfg = synth.DATASETS["kitti"]
    gs = synth.grid_size(cfg)
    net, state = _load_backbone("VoxelResBackBone8x", 4, gs, 9)
    hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000)
    frames = [synth.lidar_frame("kitti", seed=40 + i, az_steps=120 + 40 * i) for i in range(3)]
    bd, info = hp(frames, fetch="encoded")
    enc = bd["encoded_spconv_tensor"]
    feats_all, ind_all = enc.features.cpu().numpy().copy(), enc.indices.cpu().numpy().copy()
    assert np.array_equal(info["encoded_indices_host"].numpy(), ind_all)
    assert info["h2d_bytes"] > 0 and info["d2h_bytes"] > feats_all.nbytes
    for b, f in enumerate(frames):
        bd1, _ = hp([f])
        e1 = bd1["encoded_spconv_tensor"]
        sel = ind_all[:, 0] == b
        assert np.array_equal(ind_all[sel][:, 1:], e1.indices.cpu().numpy()[:, 1:])
        assert rel_err(feats_all[sel], e1.features.cpu().numpy()) < 1e-6
        # and against the CPU oracle end to end
        v, c, n = O.voxelize(f, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000)
        ref = O.backbone_forward("VoxelResBackBone8x", state, O.mean_vfe(v, n), O.collate([c]), 1,
                                 [int(gs[2]) + 1, int(gs[1]), int(gs[0])])
        assert np.array_equal(e1.indices.cpu().numpy(), ref["out"][1])
        assert rel_err(e1.features.cpu().numpy(), ref["out"][0]) < FP32_TOL


def test_bf16_backbone_within_stated_tolerance():
    g = load_golden("backbone_kitti_VoxelResBackBone8x")
    net, _ = _load_backbone("VoxelResBackBone8x", 4, g["grid_size"], int(g["seed"]), {"PRECISION": "bf16"})
    with torch.no_grad():
        bd = net({"voxel_features": cuda(g["voxel_features"]), "voxel_coords": cuda(g["voxel_coords"]),
                  "batch_size": int(g["batch_size"])})
    outs = dict(bd["multi_scale_3d_features"], out=bd["encoded_spconv_tensor"])
    for k in ("x_conv1", "x_conv2", "x_conv3", "x_conv4", "out"):
        assert outs[k].features.dtype == torch.bfloat16
        assert np.array_equal(outs[k].indices.cpu().numpy(), g[k + "_indices"]), k
        assert rel_err(outs[k].features.float().cpu().numpy(), g[k + "_features"]) < BF16_TOL, k


def test_dense_matches_reference_scatter():
    ind = synth.random_voxels([5, 12, 11], 300, 3, seed=6)
    feats = torch.randn(ind.shape[0], 16, device=DEV)
    x = spconv.SparseConvTensor(feats, cuda(ind), [5, 12, 11], 3)
    ref = spconv.scatter_nd(x.indices.long(), feats, [3, 5, 12, 11, 16]).permute(0, 4, 1, 2, 3).contiguous()
    assert torch.equal(x.dense(), ref)


def _digest(masks, kvol=27, kx=3):
    """The 12-bit grouping key of csrc/sort.cu: [x-offsets that occur | (z,y) lines with a neighbour]."""
    lines = kvol // kx
    x_bits = np.zeros_like(masks)
    line_bits = np.zeros_like(masks)
    for l in range(lines):
        seg = (masks >> (l * kx)) & ((1 << kx) - 1)
        x_bits |= seg
        line_bits |= (seg != 0).astype(masks.dtype) << l
    return (x_bits << lines) | line_bits


def test_mask_sorted_row_order_gives_identical_results():
    """fv2p_group_rows / fv2p_sort_rows_by_mask only change which rows share a tile: the conv output must be
    bit-identical.  The order is a grouping by the mask digest (ascending digest, any order inside a group)."""
    rng = np.random.default_rng(3)
    shape = [9, 40, 40]
    ind = synth.random_voxels(shape, 3000, 2, seed=5)
    ind[:, 1:] = ind[:, 1:] // np.array([2, 3, 3])
    ind = np.unique(ind, axis=0).astype(np.int32)
    for subm, st in ((True, 1), (False, 2)):
        outids, pairs, num, nbr = spconv.ops.get_indice_pairs(cuda(ind), 2, shape, 3, st, 1, 1, 0, subm, False,
                                                              return_nbr=True)
        nbr = nbr.contiguous()
        n_out = outids.shape[0]
        perm, nbr_sorted, order = spconv.ops.sort_rows_by_mask(nbr, n_out, return_tile_order=True)
        m = nbr.cpu().numpy()
        masks = ((m >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
        got = perm.cpu().numpy()
        assert np.array_equal(np.sort(got), np.arange(n_out))  # a permutation of the rows
        dg = _digest(masks)[got]
        assert np.all(np.diff(dg) >= 0)  # grouped: ascending digest
        assert np.array_equal(nbr_sorted.cpu().numpy(), m[:, got])
        # tile list: every 128-row tile once with the OR of its rows' masks, by descending number of active offsets,
        # ties in tile order
        sm = masks[got]
        tmask = np.array([int(np.bitwise_or.reduce(sm[t:t + 128])) for t in range(0, n_out, 128)])
        weight = np.array([bin(m_).count("1") for m_ in tmask])
        by_weight = np.argsort(-weight, kind="stable")
        assert np.array_equal(order.cpu().numpy()[:, 0], by_weight)
        assert np.array_equal(order.cpu().numpy()[:, 1], tmask[by_weight])
        for mode, dt in ((_lib.MODE_TF32X3_TC, torch.float32), (_lib.MODE_BF16_TC, torch.bfloat16)):
            feats = cuda(rng.standard_normal((ind.shape[0], 64)).astype(np.float32), dt)
            w = cuda((rng.standard_normal((27, 64, 64)) / 40).astype(np.float32))
            packed = spconv.ops.pack_weight(w, mode)
            res = cuda(rng.standard_normal((n_out, 64)).astype(np.float32), dt)
            a = spconv.ops.conv_forward(feats, packed, nbr, n_out, residual=res, relu=True, mode=mode)
            for kw in (dict(), dict(tile_order=order.contiguous()), dict(dynamic=False)):
                b = spconv.ops.conv_forward(feats, packed, nbr_sorted.contiguous(), n_out, residual=res, relu=True,
                                            mode=mode, row_perm=perm.contiguous(), **kw)
                assert torch.equal(a, b)
            c = spconv.ops.conv_forward(feats, packed, nbr, n_out, residual=res, relu=True, mode=mode, dynamic=False)
            assert torch.equal(a, c)


def test_run_stream_matches_single_calls_and_graph_replay():
    """The pipelined / CUDA-graph form of the hot path returns exactly what the plain call returns, batch after
    batch with different point counts (the graph is captured once for the staging capacity)."""
    cfg = synth.DATASETS["kitti"]
    net, _ = _load_backbone("VoxelResBackBone8x", 4, synth.grid_size(cfg), 9)
    batches = [[synth.lidar_frame("kitti", seed=70 + 3 * i + j, az_steps=100 + 30 * ((i + j) % 3)) for j in range(2)]
               for i in range(5)]
    ref = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000)
    expect = []
    for b in batches:
        bd, info = ref(b, fetch="encoded")
        expect.append((info["counts"], info["encoded_features_host"].clone(), info["encoded_indices_host"].clone()))
    for use_graph in (False, True):
        hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000, use_graph=use_graph)
        # size the staging for the largest batch first (a graph serves anything up to its capture capacity)
        hp.upload(max(batches, key=lambda b: sum(f.shape[0] for f in b)), slot=0)
        hp.upload(max(batches, key=lambda b: sum(f.shape[0] for f in b)), slot=1)
        hp.upload(max(batches, key=lambda b: sum(f.shape[0] for f in b)), slot=2)
        got = []
        for res in hp.run_stream(iter(batches)):
            got.append((res["counts"], res["encoded_features"].clone(), res["encoded_indices"].clone()))
        for res in hp.run_stream(iter(batches[:3]), depth=2):  # the double-buffered form
            got.append((res["counts"], res["encoded_features"].clone(), res["encoded_indices"].clone()))
        expect = expect + expect[:3] if len(expect) == len(batches) else expect
        assert len(got) == len(expect)
        for (c0, f0, i0), (c1, f1, i1) in zip(expect, got):
            assert c0 == c1
            assert torch.equal(i0, i1)
            assert torch.equal(f0, f1)


def test_engine_options_do_not_change_results():
    """MATERIALIZE_PAIRS / SORT_ROWS only change what is materialised and how rows are grouped into tiles."""
    g = load_golden("backbone_kitti_VoxelResBackBone8x")
    outs = []
    for cfg in ({}, {"MATERIALIZE_PAIRS": False}, {"SORT_ROWS": False}):
        net, _ = _load_backbone("VoxelResBackBone8x", 4, g["grid_size"], int(g["seed"]), cfg)
        with torch.no_grad():
            bd = net({"voxel_features": cuda(g["voxel_features"]), "voxel_coords": cuda(g["voxel_coords"]),
                      "batch_size": int(g["batch_size"])})
        enc = bd["encoded_spconv_tensor"]
        outs.append((enc.features.clone(), enc.indices.clone(), len(enc.indice_dict)))
    assert outs[0][2] == 9 and outs[1][2] == 0  # pair tensors are only there when asked for
    for o in outs[1:]:
        assert torch.equal(o[1], outs[0][1])
        assert torch.equal(o[0], outs[0][0])


def test_hot_path_32_frames_and_an_empty_frame():
    """One call over 32 frames (the reference's int32 dense grid stops at 23) with an empty frame in
    the middle equals the frames run one by one."""
    cfg = synth.DATASETS["kitti"]
    net, _ = _load_backbone("VoxelBackBone8x", 4, synth.grid_size(cfg), 4)
    hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000)
    frames = [synth.lidar_frame("kitti", seed=200 + i, az_steps=30 + (i % 5) * 6) for i in range(32)]
    frames[7] = np.zeros((0, 4), np.float32)
    bd, info = hp(frames)
    enc = bd["encoded_spconv_tensor"]
    ind, feat = enc.indices.cpu().numpy().copy(), enc.features.cpu().numpy().copy()
    assert not (ind[:, 0] == 7).any() and ind[:, 0].max() == 31
    assert np.all(np.diff(ind[:, 0]) >= 0)  # rows stay batch-contiguous (consumers rely on it)
    for b in (0, 6, 8, 31):
        bd1, _ = hp([frames[b]])
        e1 = bd1["encoded_spconv_tensor"]
        sel = ind[:, 0] == b
        assert np.array_equal(ind[sel][:, 1:], e1.indices.cpu().numpy()[:, 1:])
        assert rel_err(feat[sel], e1.features.cpu().numpy()) < 1e-6


def test_module_path_trains():
    """Training mode uses the plain module graph; the conv backward (a torch-index composition, not on the hot
    path) must agree with autograd through the equivalent dense conv3d."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    shape = [5, 8, 7]
    ind = synth.random_voxels(shape, 120, 1, seed=2)
    conv = spconv.SubMConv3d(6, 10, 3, bias=True, indice_key="t").to(DEV)
    feats = torch.randn(ind.shape[0], 6, device=DEV, requires_grad=True)
    x = spconv.SparseConvTensor(feats, cuda(ind), shape, 1)
    y = conv(x)
    loss = (y.features ** 2).sum()
    loss.backward()
    gw, gb, gx = conv.weight.grad.clone(), conv.bias.grad.clone(), feats.grad.clone()
    # dense reference
    w = conv.weight.detach().clone().requires_grad_(True)
    b = conv.bias.detach().clone().requires_grad_(True)
    f2 = feats.detach().clone().requires_grad_(True)
    dense = spconv.scatter_nd(cuda(ind).long(), f2, [1] + shape + [6]).permute(0, 4, 1, 2, 3)
    out = torch.nn.functional.conv3d(dense, w.permute(4, 3, 0, 1, 2), b, 1, 1)
    idx = cuda(ind).long()
    picked = out[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    ((picked ** 2).sum()).backward()
    assert torch.allclose(gw, w.grad, atol=1e-3, rtol=1e-3)
    assert torch.allclose(gb, b.grad, atol=1e-3, rtol=1e-3)
    assert torch.allclose(gx, f2.grad, atol=1e-3, rtol=1e-3)
    # strided conv: grad_input runs through fv2p_conv_fwd on the input-major map
    conv2 = spconv.SparseConv3d(6, 10, 3, stride=2, padding=1, bias=False, indice_key="t2").to(DEV)
    feats2 = torch.randn(ind.shape[0], 6, device=DEV, requires_grad=True)
    y2 = conv2(spconv.SparseConvTensor(feats2, cuda(ind), shape, 1))
    (y2.features ** 2).sum().backward()
    w2 = conv2.weight.detach().clone().requires_grad_(True)
    f3 = feats2.detach().clone().requires_grad_(True)
    dense2 = spconv.scatter_nd(cuda(ind).long(), f3, [1] + shape + [6]).permute(0, 4, 1, 2, 3)
    out2 = torch.nn.functional.conv3d(dense2, w2.permute(4, 3, 0, 1, 2), None, 2, 1)
    oi = y2.indices.long()
    picked2 = out2[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]]
    assert torch.allclose(y2.features, picked2, atol=1e-4, rtol=1e-4)
    (picked2 ** 2).sum().backward()
    assert torch.allclose(conv2.weight.grad, w2.grad, atol=1e-3, rtol=1e-3)
    assert torch.allclose(feats2.grad, f3.grad, atol=1e-3, rtol=1e-3)


def test_height_compression_matches_reference_and_oracle():
    """HeightCompression is pure data movement: bit-exact against the reference fixture, against
    the oracle at the KITTI BEV shape, in fp32 and bf16, as a module and appended to the HotPath step."""
    g = load_golden("height_compression")
    shape, batch = [int(v) for v in g["spatial_shape"]], int(g["batch_size"])
    x = spconv.SparseConvTensor(cuda(g["features"]), cuda(g["indices"]), shape, batch)
    hc = fv2p_b200.HeightCompression({"NUM_BEV_FEATURES": g["features"].shape[1] * shape[0]})
    bd = hc({"encoded_spconv_tensor": x, "encoded_spconv_tensor_stride": 8})
    assert bd["spatial_features_stride"] == 8
    assert np.array_equal(bd["spatial_features"].cpu().numpy(), g["spatial_features"])
    # KITTI stride-8 shape, fp32 and bf16, against the oracle
    rng = np.random.default_rng(21)
    shape, batch, c = [2, 200, 176], 2, 128
    ind = synth.random_voxels(shape, 5000, batch, seed=22).astype(np.int32)
    feats = rng.standard_normal((ind.shape[0], c)).astype(np.float32)
    want = O.height_compression(feats, ind, shape, batch)
    for dt in (torch.float32, torch.bfloat16):
        f = cuda(feats, dt)
        got = fv2p_b200.height_compression.height_compression(f, cuda(ind), shape, batch)
        assert got.shape == (batch, c * shape[0], shape[1], shape[2]) and got.dtype == dt
        ref_dt = torch.from_numpy(want).to(dt)
        assert torch.equal(got.cpu(), ref_dt)
    # appended to the step (row count read on the device), eager and graph replay
    cfg = synth.DATASETS["kitti"]
    net, _ = _load_backbone("VoxelResBackBone8x", 4, synth.grid_size(cfg), 4)
    frames = [synth.lidar_frame("kitti", seed=300 + i, az_steps=40) for i in range(2)]
    for use_graph in (False, True):
        hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000, use_graph=use_graph,
                               bev=True)
        for _ in range(2):
            bd, info = hp(frames)
        enc = bd["encoded_spconv_tensor"]
        want = O.height_compression(enc.features.cpu().numpy(), enc.indices.cpu().numpy(), enc.spatial_shape, 2)
        assert bd["spatial_features_stride"] == 8
        assert np.array_equal(bd["spatial_features"].cpu().numpy(), want)


def test_capacity_overflow_grows_the_arena_and_reruns():
    """The arena starts from modest row bounds (CAP_GROWTH); a step that overflows them is detected from the status word,
    the bounds grow and the step runs again - same result as with the hard bounds, eager and from a graph."""
    g = load_golden("backbone_kitti_VoxelResBackBone8x")
    want = None
    for growth in (None, 0.02):
        net, _ = _load_backbone("VoxelResBackBone8x", 4, g["grid_size"], int(g["seed"]), {"CAP_GROWTH": growth})
        with torch.no_grad():
            bd = net({"voxel_features": cuda(g["voxel_features"]), "voxel_coords": cuda(g["voxel_coords"]),
                      "batch_size": int(g["batch_size"])})
        enc = bd["encoded_spconv_tensor"]
        got = (enc.indices.cpu().numpy().copy(), enc.features.cpu().numpy().copy())
        if want is None:
            want = got
        else:
            assert net.get_engine().cap_growth is None or net.get_engine().cap_growth > 0.02  # it had to grow
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    cfg = synth.DATASETS["kitti"]
    frames = [synth.lidar_frame("kitti", seed=400 + i, az_steps=40) for i in range(2)]
    res = []
    for growth, use_graph in ((None, False), (0.02, False), (0.02, True)):
        net, _ = _load_backbone("VoxelResBackBone8x", 4, synth.grid_size(cfg), 4, {"CAP_GROWTH": growth})
        hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000, use_graph=use_graph)
        bd, info = hp(frames)
        enc = bd["encoded_spconv_tensor"]
        res.append((info["counts"], enc.indices.cpu().numpy().copy(), enc.features.cpu().numpy().copy()))
    for r in res[1:]:
        assert r[0] == res[0][0] and np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2])


def test_voxelizer_builds_the_level0_coordinate_table():
    """fv2p_voxelize_mean_table: the table the voxelizer fills while it assigns rows serves the submanifold probe
    exactly like the one fv2p_table_build makes from the finished coordinates - ragged batch, an empty frame, and a
    frame that hits max_voxels (rows past the cut must not be in the table)."""
    cfg = synth.DATASETS["kitti"]
    frames = [synth.lidar_frame("kitti", seed=s, az_steps=a) for s, a in ((11, 200), (12, 60), (13, 500))]
    frames.insert(1, np.zeros((0, 4), np.float32))
    max_voxels = 7000  # the last frame has more voxels than that
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int32)
    lib = _lib.load()
    shape = [int(v) for v in (np.array(synth.grid_size(cfg))[::-1] + [1, 0, 0])]
    bv = fv2p_b200.BatchVoxelizer(cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels)
    pts = cuda(np.concatenate(frames))
    cap = bv.capacity(pts.device, pts.shape[0], len(frames), 4)
    row_cap = cap + 1000
    table = torch.empty(lib.fv2p_table_bytes(row_cap) + 16, dtype=torch.uint8, device="cuda")
    table.fill_(0x5A)  # garbage: the call clears it
    out = bv(pts, cuda(offs), max(f.shape[0] for f in frames), level0_table=(table, row_cap, shape))
    voff = out["voxel_offsets"].cpu().numpy()
    m = int(voff[-1])
    assert voff[2] == voff[1] and voff[4] - voff[3] == max_voxels
    coords = out["voxel_coords"][:m].contiguous()
    # the plain call gives the same rows (and the oracle checks of the other tests apply to it)
    plain = fv2p_b200.BatchVoxelizer(cfg["voxel_size"], cfg["point_cloud_range"], 5, max_voxels)(
        pts, cuda(offs), max(f.shape[0] for f in frames))
    assert np.array_equal(plain["voxel_offsets"].cpu().numpy(), voff)
    assert torch.equal(plain["voxel_coords"][:m], coords)
    assert torch.equal(plain["voxel_features"][:m], out["voxel_features"][:m])
    ref_table = torch.zeros_like(table)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.fv2p_table_build(_lib.ptr(coords), m, None, _lib.i32x3(shape), _lib.ptr(ref_table), row_cap,
                                    _lib.ptr(status), 0, _lib.stream_ptr(coords.device)), "table_build")
    nbrs = []
    for t in (table, ref_table):
        nbr = torch.empty((27, (m + 127) // 128 * 128), dtype=torch.int32, device="cuda")
        _lib.check(lib.fv2p_subm_neighbours(_lib.ptr(coords), m, None, len(frames), _lib.i32x3(shape),
                                            _lib.i32x3([3, 3, 3]), _lib.i32x3([1, 1, 1]), _lib.ptr(t), row_cap,
                                            _lib.ptr(nbr), nbr.shape[1], 0, _lib.stream_ptr(coords.device)),
                   "subm_neighbours")
        nbrs.append(nbr[:, :m])
    assert torch.equal(nbrs[0], nbrs[1])
    assert int((nbrs[0][13] == torch.arange(m, device="cuda", dtype=torch.int32)).all())
    # slot-for-slot: same keys present with the same rows (probe order may place them in different slots)
    a = np.sort(table.cpu().numpy()[:lib.fv2p_table_bytes(row_cap)].view(np.int64).reshape(-1, 2), axis=0)
    b = np.sort(ref_table.cpu().numpy()[:lib.fv2p_table_bytes(row_cap)].view(np.int64).reshape(-1, 2), axis=0)
    assert np.array_equal(a, b)


def test_tma_gather_and_cp_async_gather_give_identical_results():
    """The two producers of the gathered A tile (fv2p_tc_gather_mode: TMA tile::gather4 = the default for stages of one
    offset, swizzled cp.async) only differ in how the rows reach shared memory: bit-identical outputs, grouped and
    ungrouped, including the zero fill of missing neighbours and a partial last tile."""
    rng = np.random.default_rng(11)
    shape = [9, 40, 40]
    ind = synth.random_voxels(shape, 3001, 2, seed=6)
    lib = _lib.load()
    outids, pairs, num, nbr = spconv.ops.get_indice_pairs(cuda(ind), 2, shape, 3, 1, 1, 1, 0, True, False,
                                                          return_nbr=True)
    nbr = nbr.contiguous()
    n_out = outids.shape[0]
    perm, nbr_sorted, order = spconv.ops.sort_rows_by_mask(nbr, n_out, return_tile_order=True)
    try:
        for mode, dt, shapes in ((_lib.MODE_FP32_TC, torch.float32, ((32, 32), (64, 64), (128, 128), (64, 128))),
                                 (_lib.MODE_BF16_TC, torch.bfloat16, ((64, 64), (128, 128), (64, 128)))):
            for cin, cout in shapes:
                feats = cuda(rng.standard_normal((ind.shape[0], cin)).astype(np.float32), dt)
                packed = spconv.ops.pack_weight(cuda((rng.standard_normal((27, cin, cout)) / 40).astype(np.float32)), mode)
                outs = []
                for g in (0, 1):
                    lib.fv2p_tc_gather_mode(g)
                    outs.append(spconv.ops.conv_forward(feats, packed, nbr, n_out, relu=False, mode=mode))
                    outs.append(spconv.ops.conv_forward(feats, packed, nbr_sorted.contiguous(), n_out, relu=False,
                                                        mode=mode, row_perm=perm.contiguous(),
                                                        tile_order=order.contiguous()))
                assert float(outs[0].float().abs().max()) > 0
                for o in outs[1:]:
                    assert torch.equal(outs[0], o)
    finally:
        lib.fv2p_tc_gather_mode(-1)


def test_run_stream_cold_start_with_a_busy_default_stream():
    """The staging buffers of a fresh HotPath are created while run_stream is already enqueueing copies on its own
    streams: nothing may depend on work queued on the default stream (a zero-filled frame-offset buffer once did - its
    fill kernel ran on the current stream, unordered with the H2D copy on the copy stream, and a batch that lost the
    race saw no points; found under compute-sanitizer, where the timing differs)."""
    cfg = synth.DATASETS["kitti"]
    net, _ = _load_backbone("VoxelResBackBone8x", 4, synth.grid_size(cfg), 9)
    batches = [[synth.lidar_frame("kitti", seed=40 + 2 * i + j, az_steps=90) for j in range(2)] for i in range(4)]
    ref = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000)
    expect = [ref(b, fetch="encoded")[1]["counts"] for b in batches]
    for use_graph in (False, True):
        hp = fv2p_b200.HotPath(net, cfg["voxel_size"], cfg["point_cloud_range"], 5, 16000, use_graph=use_graph)
        ballast = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        for _ in range(40):  # tens of milliseconds of queued work on the default stream
            ballast.zero_()
        got = [res["counts"] for res in hp.run_stream(iter(batches))]
        torch.cuda.synchronize()
        assert got == expect
        assert all(c[-1] > 0 for c in got)
