"""voxel -> point 3-NN interpolation (front end of ResidualVoxelToPointDecoder) and voxel_query on
the level's coordinate table.

GPU tests compare libfv2p_b200 with the REFERENCE's own CUDA kernels (three_nn / three_interpolate /
voxel_query_kernel_stack compiled from /root/reference into oracle/_ref/libref_pointnet2.so, oracle/build_ref.py) -
indices and distances bit-exact, interpolated features to 1e-6 - and with the numpy oracle (which is thereby pinned
against the reference on the same inputs).  CPU tests check the oracle's own logic on small cases.
"""
import numpy as np
import pytest
import torch

import fv2p_b200
from fv2p_b200 import pointops, spconv, synth
from conftest import rel_err
from oracle import oracle as O
from oracle import ref as R

DEV = "cuda:0"
VOXEL_SIZE, PC_RANGE, DS = [0.05, 0.05, 0.1], [0.0, -40.0, -3.0, 70.4, 40.0, 1.0], 8
SHAPE = [5, 200, 176]  # stride-8 level of the KITTI grid


def cuda(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _scene(n_vox, n_pts, batch, seed, channels=32):
    rng = np.random.default_rng(seed)
    ind = synth.random_voxels(SHAPE, n_vox, batch, seed=seed)
    ind = ind[np.lexsort((ind[:, 3], ind[:, 2], ind[:, 1], ind[:, 0]))]  # any order works; frames stay contiguous
    ind = ind[rng.permutation(ind.shape[0])]
    ind = ind[np.argsort(ind[:, 0], kind="stable")].astype(np.int32)
    feats = rng.standard_normal((ind.shape[0], channels)).astype(np.float32)
    pts = np.zeros((n_pts, 4), np.float32)
    pts[:, 0] = rng.integers(0, batch, n_pts)
    pts[:, 1] = rng.uniform(0.0, 70.4, n_pts)
    pts[:, 2] = rng.uniform(-40.0, 40.0, n_pts)
    pts[:, 3] = rng.uniform(-3.0, 1.0, n_pts)
    pts[: n_pts // 50, 1:] += rng.uniform(-30.0, 30.0, (n_pts // 50, 3)).astype(np.float32)  # some outside the range
    pts = pts[np.argsort(pts[:, 0], kind="stable")]
    return ind, feats, pts


# ------------------------------------------------------------------------------------------------ CPU: oracle logic
def test_oracle_three_nn_against_a_plain_loop():
    rng = np.random.default_rng(1)
    known = rng.random((40, 3)).astype(np.float32)
    known[7] = known[3]  # an exact tie: the lower index must win
    unknown = np.concatenate([rng.random((9, 3)).astype(np.float32), known[3:4]])
    dist, idx, d2 = O.three_nn(unknown, known)
    for u in range(unknown.shape[0]):
        best = [(1e40, 0)] * 3
        for k in range(known.shape[0]):
            diff = (unknown[u] - known[k]).astype(np.float32).astype(np.float64)
            d = float((diff * diff).sum())
            if d < best[0][0]:
                best = [(d, k), best[0], best[1]]
            elif d < best[1][0]:
                best = [best[0], (d, k), best[1]]
            elif d < best[2][0]:
                best = [best[0], best[1], (d, k)]
        assert [b[1] for b in best] == idx[u].tolist()
    assert idx[-1, 0] == 3 and idx[-1, 1] == 7 and dist[-1, 0] == 0.0
    few = O.three_nn(unknown, known[:2])
    assert np.isinf(few[0][:, 2]).all() and (few[1][:, 2] == 0).all()


def test_oracle_voxel_query_first_nsample_rule_and_empty_ball():
    ind = np.int32([[0, 1, 1, 1], [0, 1, 1, 2], [0, 1, 2, 1], [0, 3, 3, 3]])
    grid = O.voxel2pinds(ind, [4, 4, 4], 1)
    assert grid[0, 1, 1, 2] == 1 and grid[0, 0, 0, 0] == -1
    xyz = np.float32([[1, 1, 1], [2, 1, 1], [1, 2, 1], [3, 3, 3]])
    new_xyz = np.float32([[1.1, 1.0, 1.0], [0.0, 3.0, 0.0]])
    new_coords = np.int32([[0, 1, 1, 1], [0, 0, 3, 0]])
    idx, empty = O.voxel_query((1, 1, 1), 1.05, 4, xyz, new_xyz, new_coords, grid)
    assert idx[0].tolist() == [0, 1, 2, 0] and not empty[0]  # visiting order z, y, x; unfilled slots repeat the first
    assert empty[1] and idx[1].tolist() == [0, 0, 0, 0]


# ------------------------------------------------------------------------------------------------ GPU
needs_ref = pytest.mark.skipif(not R.have_pointnet2(), reason="oracle/_ref/libref_pointnet2.so is not built")


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("n_vox,n_pts,batch", [(6000, 20000, 2), (300, 5000, 3), (40000, 60000, 4)])
def test_voxel_three_nn_bit_exact_against_reference_kernels(n_vox, n_pts, batch):
    ind, feats, pts = _scene(n_vox, n_pts, batch, seed=n_vox)
    x = spconv.SparseConvTensor(cuda(feats), cuda(ind), SHAPE, batch)
    dist, idx, out = pointops.voxel_three_nn(x, cuda(pts), VOXEL_SIZE, PC_RANGE, DS, features=cuda(feats))
    centres = pointops.get_voxel_centers(cuda(ind)[:, 1:4], DS, VOXEL_SIZE, PC_RANGE)
    for b in range(batch):
        vm, pm = cuda(ind)[:, 0] == b, cuda(pts)[:, 0].long() == b
        r_out, r_dist, r_idx = R.ref_top3_interpolate(centres[vm], cuda(pts)[pm][:, 1:4].contiguous(),
                                                      cuda(feats)[vm])
        assert torch.equal(idx[pm], r_idx), b
        assert torch.equal(dist[pm], r_dist), b
        a_, b_ = out[pm].cpu().numpy(), r_out.cpu().numpy()
        bad = np.abs(a_ - b_).max(1) > 1e-5 * max(np.abs(b_).max(), 1e-30)
        assert rel_err(a_, b_) < 1e-6, (b, int(bad.sum()), a_.shape, np.abs(a_ - b_).max(), a_[bad][:2], b_[bad][:2],
                                          dist[pm].cpu().numpy()[bad][:2], idx[pm].cpu().numpy()[bad][:2])
    # and the oracle, pinned here against the same reference outputs: same neighbours wherever the fp32 distances
    # are not within rounding of a tie, distances to 1e-6
    sel = np.sort(np.random.default_rng(0).choice(pts.shape[0], min(4000, pts.shape[0]), replace=False))
    o_out, o_dist, o_idx = O.voxel_to_point_interpolate(ind, feats, pts[sel], batch, VOXEL_SIZE, PC_RANGE, DS)
    g_idx, g_dist, g_out = idx.cpu().numpy()[sel], dist.cpu().numpy()[sel], out.cpu().numpy()[sel]
    assert np.allclose(o_dist, g_dist, rtol=2e-6, atol=1e-6)
    differ = (o_idx != g_idx).any(1)
    assert differ.mean() < 2e-3
    assert rel_err(o_out[~differ], g_out[~differ]) < 1e-5


@pytest.mark.gpu
@needs_ref
def test_voxel_three_nn_sparse_frames_and_far_points():
    """Frames with fewer than three voxels (the reference leaves distance inf / index 0), an empty frame, and points
    far from every voxel (exact-scan fallback)."""
    ind = np.int32([[0, 1, 10, 10], [0, 2, 150, 100], [2, 0, 5, 5], [2, 4, 199, 175], [2, 2, 100, 90], [2, 2, 100, 91]])
    feats = np.arange(12, dtype=np.float32).reshape(6, 2)
    pts = np.float32([[0, 1.0, -39.0, -2.5], [0, 69.0, 39.0, 0.5], [2, 35.0, 0.1, -1.0], [2, 0.1, -39.9, -2.9],
                      [2, 70.3, 39.9, 0.9]])
    x = spconv.SparseConvTensor(cuda(feats), cuda(ind), SHAPE, 3)
    dist, idx, out = pointops.voxel_three_nn(x, cuda(pts), VOXEL_SIZE, PC_RANGE, DS, features=cuda(feats))
    centres = pointops.get_voxel_centers(cuda(ind)[:, 1:4], DS, VOXEL_SIZE, PC_RANGE)
    for b in (0, 2):
        vm, pm = cuda(ind)[:, 0] == b, cuda(pts)[:, 0].long() == b
        r_out, r_dist, r_idx = R.ref_top3_interpolate(centres[vm], cuda(pts)[pm][:, 1:4].contiguous(), cuda(feats)[vm])
        assert torch.equal(idx[pm], r_idx) and torch.equal(dist[pm], r_dist)
        assert torch.allclose(out[pm], r_out, rtol=1e-6, atol=1e-6, equal_nan=True)
    assert torch.isinf(dist[:2, 2]).all() and (idx[:2, 2] == 0).all()


@pytest.mark.gpu
@needs_ref
def test_voxel_three_nn_on_backbone_outputs():
    """The decoder's real inputs: x_conv3 / x_conv4 of a KITTI frame pair from HotPath (the tables come from the
    engine), raw points as queries."""
    cfg = synth.DATASETS["kitti"]
    gs = synth.grid_size(cfg)
    frames = [synth.lidar_frame("kitti", seed=50 + i) for i in range(2)]
    net = fv2p_b200.VoxelResBackBone8x({}, 4, np.array(gs)).eval()
    state = synth.randomize_state(net.state_dict(), seed=3)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=False)
    hp = fv2p_b200.HotPath(net.to(DEV), cfg["voxel_size"], cfg["point_cloud_range"], 5, 40000)
    bd, _ = hp(frames, device=DEV)
    pts = np.concatenate([np.concatenate([np.full((f.shape[0], 1), b, np.float32), f[:, :3]], 1)
                          for b, f in enumerate(frames)])
    for name, ds in (("x_conv3", 4), ("x_conv4", 8)):
        sp = bd["multi_scale_3d_features"][name]
        assert getattr(sp, "fv2p_table", None) is not None
        got = pointops.voxel_to_point_interpolate(sp, cuda(pts), cfg["voxel_size"], cfg["point_cloud_range"], ds)
        centres = pointops.get_voxel_centers(sp.indices[:, 1:4], ds, cfg["voxel_size"], cfg["point_cloud_range"])
        for b in range(2):
            vm, pm = sp.indices[:, 0] == b, cuda(pts)[:, 0].long() == b
            r_out, _, _ = R.ref_top3_interpolate(centres[vm], cuda(pts)[pm][:, 1:4].contiguous(),
                                                 sp.features[vm].float())
            assert rel_err(got[pm].cpu().numpy(), r_out.cpu().numpy()) < 1e-6, (name, b)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("max_range,radius,nsample", [((1, 1, 1), 0.6, 16), ((2, 3, 3), 1.2, 8), ((0, 2, 2), 5.0, 4)])
def test_voxel_query_bit_exact_against_reference_kernel(max_range, radius, nsample):
    ind, feats, pts = _scene(8000, 4000, 2, seed=17)
    x = spconv.SparseConvTensor(cuda(feats), cuda(ind), SHAPE, 2)
    xyz = pointops.get_voxel_centers(cuda(ind)[:, 1:4], DS, VOXEL_SIZE, PC_RANGE).contiguous()
    # query centres: points inside the range, with the voxel they fall into (voxelrcnn_head.py:147-175)
    p = pts[(pts[:, 1] > 0) & (pts[:, 1] < 70.4) & (pts[:, 2] > -40) & (pts[:, 2] < 40) & (pts[:, 3] > -3) &
            (pts[:, 3] < 1)]
    vs = np.float32(VOXEL_SIZE) * DS
    cz = np.floor((p[:, 3] + 3.0) / vs[2]).astype(np.int32)
    cy = np.floor((p[:, 2] + 40.0) / vs[1]).astype(np.int32)
    cx = np.floor(p[:, 1] / vs[0]).astype(np.int32)
    new_coords = np.stack([p[:, 0].astype(np.int32), cz, cy, cx], 1).astype(np.int32)
    new_xyz = np.ascontiguousarray(p[:, 1:4])
    v2p = pointops.generate_voxel2pinds(x)
    idx, empty = pointops.voxel_query(max_range, radius, nsample, xyz, cuda(new_xyz), cuda(new_coords), v2p)
    dense = v2p.dense()
    assert np.array_equal(dense.cpu().numpy(), O.voxel2pinds(ind, SHAPE, 2))
    r_idx, r_empty = R.ref_voxel_query(max_range, radius, nsample, xyz, cuda(new_xyz), cuda(new_coords), dense)
    assert torch.equal(idx, r_idx) and torch.equal(empty, r_empty)
    assert 0 < int(empty.sum()) < empty.numel() or radius > 1.0
    o_idx, o_empty = O.voxel_query(max_range, radius, nsample, xyz.cpu().numpy(), new_xyz[:300], new_coords[:300],
                                   dense.cpu().numpy())
    same = (o_idx == idx.cpu().numpy()[:300]).all(1)
    assert same.mean() > 0.995  # fp32-vs-float64 distance at the radius threshold
    with pytest.raises(TypeError):
        pointops.voxel_query(max_range, radius, nsample, xyz, cuda(new_xyz), cuda(new_coords), dense)
