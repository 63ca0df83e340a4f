"""Consumers of the backbone's sparse outputs, on the level's coordinate table.

Mirrors, name for name, what the detectors call on the reference:

  * ``get_voxel_centers`` (pcdet/utils/common_utils.py:76-92),
  * ``three_nn`` / ``top3_interpolate`` as ``ResidualVoxelToPointDecoder.forward`` uses them on voxel centres
    (pcdet/models/backbones_3d/pfe/residual_v2p_decoder.py:86-116; pcdet/ops/pointnet2/pointnet2_batch/
    pointnet2_utils.py:292-326), fused into ``voxel_to_point_interpolate`` - one call for the whole batch instead of a
    python loop over frames with a brute-force O(P*N) search each,
  * ``generate_voxel2pinds`` (pcdet/utils/spconv_utils.py:13-21) and ``voxel_query``
    (pcdet/ops/pointnet2/pointnet2_stack/voxel_query_utils.py:10-48): the dense [B,Z,Y,X] index grid becomes a
    ``VoxelIndexTable`` (the O(N) hash table the rulebooks already use).

Everything runs through libfv2p_b200.so; there is no CPU or PyTorch fallback.
"""
import numpy as np
import torch

from . import _lib


def get_voxel_centers(voxel_coords, downsample_times, voxel_size, point_cloud_range):
    """common_utils.py:76-92: voxel_coords [N,3] (z,y,x) -> centres [N,3] (x,y,z), fp32."""
    assert voxel_coords.shape[1] == 3
    voxel_centers = voxel_coords[:, [2, 1, 0]].float()
    voxel_size = torch.tensor(voxel_size, device=voxel_centers.device).float() * downsample_times
    pc_range = torch.tensor(point_cloud_range[0:3], device=voxel_centers.device).float()
    return (voxel_centers + 0.5) * voxel_size + pc_range


class VoxelIndexTable(object):
    """(batch, z, y, x) -> row of a sparse tensor: what ``generate_voxel2pinds`` returns here instead of the dense
    int32 grid.  ``dense()`` materialises the reference's tensor for callers that really want it."""

    def __init__(self, table, row_cap, indices, spatial_shape, batch_size):
        self.table, self.row_cap = table, int(row_cap)
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)

    @property
    def shape(self):
        return tuple([self.batch_size] + self.spatial_shape)

    def dense(self):
        out = -torch.ones(self.shape, dtype=torch.int32, device=self.indices.device)
        ind = self.indices.long()
        out[ind[:, 0], ind[:, 1], ind[:, 2], ind[:, 3]] = torch.arange(ind.shape[0], device=ind.device,
                                                                      dtype=torch.int32)
        return out


def _table_of(sparse_tensor):
    """The level's table: handed over by the fused engine when the tensor came from it, else built here."""
    cached = getattr(sparse_tensor, "fv2p_table", None)
    if cached is not None:
        return cached
    ind = sparse_tensor.indices
    dev = _lib.require_device(ind)
    if ind.dtype != torch.int32:
        raise ValueError("indices must be int32")
    ind = ind.contiguous()
    lib = _lib.load()
    n = int(ind.shape[0])
    table = torch.empty(lib.fv2p_table_bytes(max(n, 1)) + 16, dtype=torch.uint8, device=ind.device)
    with torch.cuda.device(dev):
        _lib.check(lib.fv2p_table_build(_lib.ptr(ind), n, None, _lib.i32x3(sparse_tensor.spatial_shape),
                                        _lib.ptr(table), max(n, 1), None, 0, _lib.stream_ptr(ind.device)),
                   "table_build")
    sparse_tensor.fv2p_table = (table, max(n, 1))
    return sparse_tensor.fv2p_table


def generate_voxel2pinds(sparse_tensor):
    """spconv_utils.py:13-21.  Returns a VoxelIndexTable (pass it to ``voxel_query`` as ``point_indices``)."""
    table, row_cap = _table_of(sparse_tensor)
    return VoxelIndexTable(table, row_cap, sparse_tensor.indices, sparse_tensor.spatial_shape,
                           sparse_tensor.batch_size)


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """voxel_query_utils.py:10-48: same arguments, ``point_indices`` being the VoxelIndexTable of
    ``generate_voxel2pinds``.  Returns (idx [M,nsample] int32, empty_ball_mask [M] bool)."""
    if not isinstance(point_indices, VoxelIndexTable):
        raise TypeError("voxel_query: point_indices must come from fv2p_b200.pointops.generate_voxel2pinds (the dense "
                        "[B,Z,Y,X] grid of the reference is not built here)")
    dev = _lib.require_device(new_xyz)
    assert new_xyz.is_contiguous() and xyz.is_contiguous() and new_coords.is_contiguous()
    if new_coords.dtype != torch.int32 or xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise ValueError("voxel_query expects int32 coordinates and float32 positions")
    m = int(new_coords.shape[0])
    idx = torch.empty((m, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    z_range, y_range, x_range = max_range
    with torch.cuda.device(dev):
        st = _lib.load().fv2p_voxel_query(m, _lib.i32x3(point_indices.spatial_shape), int(nsample), float(radius),
                                          _lib.i32x3([z_range, y_range, x_range]), _lib.ptr(new_xyz), _lib.ptr(xyz),
                                          _lib.ptr(new_coords), _lib.ptr(point_indices.table), point_indices.row_cap,
                                          _lib.ptr(idx), _lib.stream_ptr(new_xyz.device))
    _lib.check(st, "voxel_query")
    empty_ball_mask = (idx[:, 0] == -1)
    idx[empty_ball_mask] = 0
    return idx, empty_ball_mask


def voxel_three_nn(sparse_tensor, point_coords, voxel_size, point_cloud_range, downsample_times, features=None):
    """For every point of ``point_coords`` [P,4] (batch, x, y, z) the three nearest voxel centres of its frame.

    Returns (dist [P,3] fp32, idx [P,3] int32 row inside the point's frame) - what the reference gets from
    ``three_nn(new_xyz, get_voxel_centers(...))`` frame by frame - and, with ``features`` [N,C], also the
    interpolated features [P,C] of ``top3_interpolate``."""
    dev = _lib.require_device(point_coords)
    if point_coords.dtype != torch.float32 or point_coords.shape[1] != 4:
        raise ValueError("point_coords must be float32 [P,4] (batch, x, y, z)")
    pts = point_coords.contiguous()
    ind = sparse_tensor.indices.contiguous()
    table, row_cap = _table_of(sparse_tensor)
    lib = _lib.load()
    p = int(pts.shape[0])
    batch = int(sparse_tensor.batch_size)
    vs = (np.asarray(voxel_size, np.float32) * np.float32(downsample_times)).astype(np.float32)
    lo = np.asarray(point_cloud_range[0:3], np.float32)
    dist = torch.empty((p, 3), dtype=torch.float32, device=pts.device)
    idx = torch.empty((p, 3), dtype=torch.int32, device=pts.device)
    out = None
    feats = None
    if features is not None:
        feats = features.contiguous()
        if feats.dtype != torch.float32 or feats.shape[0] != ind.shape[0]:
            raise ValueError("features must be float32 [N,C] aligned with the sparse tensor's rows")
        out = torch.empty((p, feats.shape[1]), dtype=torch.float32, device=pts.device)
    ws = _lib.Workspace.get(pts.device, lib.fv2p_voxel_three_nn_workspace_bytes(batch, p), "three_nn")
    with torch.cuda.device(dev):
        st = lib.fv2p_voxel_three_nn(_lib.ptr(pts), p, _lib.ptr(ind), int(ind.shape[0]), None, batch,
                                     _lib.i32x3(sparse_tensor.spatial_shape), _lib.ptr(table), row_cap,
                                     _lib.f32arr(vs), _lib.f32arr(lo), _lib.ptr(dist), _lib.ptr(idx), _lib.ptr(feats),
                                     int(feats.shape[1]) if feats is not None else 0, _lib.ptr(out), _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr(pts.device))
    _lib.check(st, "voxel_three_nn")
    if features is not None:
        return dist, idx, out
    return dist, idx


def voxel_to_point_interpolate(sparse_tensor, point_coords, voxel_size, point_cloud_range, downsample_times):
    """Steps 2-3 of ResidualVoxelToPointDecoder.forward (residual_v2p_decoder.py:94-116) for the whole batch:
    voxel centres, per-frame 3-NN, inverse-distance interpolation of ``sparse_tensor.features`` onto the points.
    Returns [P, C] fp32 aligned with ``point_coords``."""
    return voxel_three_nn(sparse_tensor, point_coords, voxel_size, point_cloud_range, downsample_times,
                          features=sparse_tensor.features.float())[2]
