"""CPU parity oracle -- numpy/ctypes front end of ``fv2p_oracle.c``.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product package never does (tests/test_boundary.py greps for it).

Parity status: pinned against the reference run in the build container, see
``tests/golden/make_golden.py`` and ``tests/test_oracle.py``.

Reference citations are relative to /root/reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfv2p_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "fv2p_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_voxelize.restype = ctypes.c_int
        _lib.orc_rulebook_subm.restype = ctypes.c_int
        _lib.orc_rulebook_conv.restype = ctypes.c_int
        _lib.orc_indice_conv.restype = ctypes.c_int
        _lib.orc_valid_out_pos.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _triple(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


# --------------------------------------------------------------------------------------------
# voxelizer + MeanVFE
# --------------------------------------------------------------------------------------------
def grid_size(voxel_size, point_cloud_range):
    """voxel_generator.py:22-27 -- fp32 arithmetic, round half to even."""
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def voxelize(points, voxel_size, point_cloud_range, max_points, max_voxels):
    """VoxelGenerator.generate (voxel_generator.py:35-39, 75-133, 136-207).

    Returns (voxels [M,T,F] f32, coors [M,3] int32 (z,y,x), num_points [M] int32).
    """
    points = _f32(points)
    P, F = points.shape
    rng = _f32(point_cloud_range)
    vs = _f32(voxel_size)
    voxels = np.empty((max_voxels, max_points, F), np.float32)
    coors = np.empty((max_voxels, 3), np.int32)
    num = np.empty((max_voxels,), np.int32)
    m = lib().orc_voxelize(_fp(points), P, F, _fp(rng), _fp(vs), int(max_points), int(max_voxels),
                           _fp(voxels), _ip(coors), _ip(num))
    if m < 0:
        raise MemoryError("oracle voxelizer could not allocate its dense lookup")
    return voxels[:m].copy(), coors[:m].copy(), num[:m].copy()


def mean_vfe(voxels, num_points):
    """MeanVFE.forward (mean_vfe.py:26-28)."""
    voxels = _f32(voxels)
    num_points = _i32(num_points)
    M, T, F = voxels.shape
    out = np.empty((M, F), np.float32)
    lib().orc_mean_vfe(_fp(voxels), _ip(num_points), M, T, F, _fp(out))
    return out


def collate(coors_list):
    """DatasetTemplate.collate_batch (dataset.py:164-169): prepend the batch index."""
    out = []
    for b, c in enumerate(coors_list):
        out.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c.astype(np.int32)], 1))
    return np.concatenate(out, 0) if out else np.zeros((0, 4), np.int32)


# --------------------------------------------------------------------------------------------
# rulebooks
# --------------------------------------------------------------------------------------------
def conv_output_size(in_shape, ksize, stride, pad, dil):
    """ops.py:20-30."""
    return [(in_shape[i] + 2 * pad[i] - dil[i] * (ksize[i] - 1) - 1) // stride[i] + 1 for i in range(3)]


def valid_out_pos(pos, ksize, stride, pad, dil, out_shape):
    """geometry.h:25-85 for one coordinate; returns (records [n,4], raw enumeration indices [n])."""
    ks, st, pd, dl, osz = (_i32(_triple(v)) for v in (ksize, stride, pad, dil, out_shape))
    kvol = int(np.prod(ks))
    rec = np.zeros((kvol, 4), np.int32)
    raw = np.zeros((kvol,), np.int32)
    n = lib().orc_valid_out_pos(_ip(_i32(pos)), _ip(ks), _ip(st), _ip(pd), _ip(dl), _ip(osz), _ip(rec), _ip(raw))
    return rec[:n].copy(), raw[:n].copy()


def rulebook_subm(indices, batch, shape, ksize=3, dilation=1):
    """get_indice_pairs(subm=True): spconv_ops.h:28-104 + geometry.h:248-297.

    Returns (outids (== indices), pairs [K,2,N] int32, num [K] int32).
    """
    indices = _i32(indices)
    n = indices.shape[0]
    ks, dl, shp = _i32(_triple(ksize)), _i32(_triple(dilation)), _i32(_triple(shape))
    kvol = int(np.prod(ks))
    pairs = np.empty((kvol, 2, n), np.int32)
    num = np.empty((kvol,), np.int32)
    r = lib().orc_rulebook_subm(_ip(indices), n, int(batch), _ip(shp), _ip(ks), _ip(dl), _ip(pairs), _ip(num))
    if r < 0:
        raise MemoryError("oracle rulebook grid allocation failed")
    return indices, pairs, num


def rulebook_conv(indices, batch, in_shape, ksize, stride, pad, dilation=1):
    """get_indice_pairs(subm=False): spconv_ops.h:28-141 + geometry.h:145-194.

    Returns (outids [Nout,4], pairs [K,2,N], num [K], out_shape).
    """
    indices = _i32(indices)
    n = indices.shape[0]
    ks, st, pd, dl = (_triple(v) for v in (ksize, stride, pad, dilation))
    out_shape = conv_output_size(_triple(in_shape), ks, st, pd, dl)
    kvol = int(np.prod(ks))
    pairs = np.empty((kvol, 2, n), np.int32)
    num = np.empty((kvol,), np.int32)
    outids = np.empty((max(n * kvol, 1), 4), np.int32)
    n_out = lib().orc_rulebook_conv(_ip(indices), n, int(batch), _ip(_i32(out_shape)), _ip(_i32(ks)),
                                    _ip(_i32(st)), _ip(_i32(pd)), _ip(_i32(dl)), _ip(outids), _ip(pairs), _ip(num))
    if n_out < 0:
        raise MemoryError("oracle rulebook grid allocation failed")
    return outids[:n_out].copy(), pairs, num, out_shape


def height_compression(features, indices, spatial_shape, batch_size):
    """HeightCompression.forward (pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-25) on top of
    SparseConvTensor.dense() (pcdet/ops/spconv/structure.py:5-18, 57-66): scatter [N,C] rows at (b,z,y,x) into zeros
    [B,D,H,W,C], permute to [B,C,D,H,W], view as [B,C*D,H,W]."""
    features = np.asarray(features)
    indices = _i32(indices)
    d, h, w = (int(v) for v in spatial_shape)
    dense = np.zeros((int(batch_size), d, h, w, features.shape[1]), features.dtype)
    dense[indices[:, 0], indices[:, 1], indices[:, 2], indices[:, 3]] = features
    return np.ascontiguousarray(dense.transpose(0, 4, 1, 2, 3)).reshape(int(batch_size), features.shape[1] * d, h, w)


# --------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------
def indice_conv(features, filters, pairs, num, n_out, inverse=False, subm=False):
    """indice_conv_fp32 (spconv_ops.h:261-362). filters may be [kD,kH,kW,Cin,Cout] or [K,Cin,Cout]."""
    features = _f32(features)
    filters = _f32(filters)
    cin, cout = filters.shape[-2], filters.shape[-1]
    filters = filters.reshape(-1, cin, cout)
    pairs = _i32(pairs)
    num = _i32(num)
    kvol = filters.shape[0]
    assert pairs.shape[0] == kvol and features.shape[1] == cin
    out = np.empty((int(n_out), cout), np.float32)
    r = lib().orc_indice_conv(_fp(features), _fp(filters), _ip(pairs), _ip(num), pairs.shape[2],
                              features.shape[0], int(n_out), kvol, cin, cout, int(inverse), int(subm), _fp(out))
    if r < 0:
        raise MemoryError
    return out


def bias_bn_res_relu(x, bias=None, bn=None, residual=None, relu=True, eps=1e-3):
    """conv.py:223-224 + BatchNorm1d(eval) + optional residual + ReLU, in place on a copy.

    bn = (gamma, beta, running_mean, running_var) or None.
    """
    x = _f32(x).copy()
    rows, ch = x.shape
    b = _f32(bias) if bias is not None else None
    g = be = mu = var = None
    if bn is not None:
        g, be, mu, var = (_f32(t) for t in bn)
    res = _f32(residual) if residual is not None else None
    lib().orc_bias_bn_res_relu(_fp(x), rows, ch, _fp(b), _fp(g), _fp(be), _fp(mu), _fp(var),
                               ctypes.c_float(eps), _fp(res), int(relu))
    return x


# --------------------------------------------------------------------------------------------
# backbones (spconv_backbone.py:71-186 VoxelBackBone8x, :189-290 VoxelResBackBone8x)
# The plan is data: every entry names the state_dict prefix of the conv and of its BatchNorm.
# --------------------------------------------------------------------------------------------
def _block(prefix, cin, cout, key, kind="subm", ksize=3, stride=1, pad=1):
    # post_act_block (spconv_backbone.py:10-29): conv(bias=False) + BN + ReLU
    return dict(op="conv", conv=prefix + ".0", bn=prefix + ".1", cin=cin, cout=cout, key=key, kind=kind,
                ksize=ksize, stride=stride, pad=pad, relu=True)


def _basic(prefix, ch, key):
    # SparseBasicBlock (spconv_backbone.py:32-68)
    return dict(op="basic", prefix=prefix, ch=ch, key=key)


def backbone_plan(name, in_ch, last_pad=0):
    if name == "VoxelBackBone8x":
        c4 = 64
        stages = [
            ("conv_input", [dict(op="conv", conv="conv_input.0", bn="conv_input.1", cin=in_ch, cout=16, key="subm1",
                                 kind="subm", ksize=3, stride=1, pad=1, relu=True)]),
            ("conv1", [_block("conv1.0", 16, 16, "subm1")]),
            ("conv2", [_block("conv2.0", 16, 32, "spconv2", "spconv", 3, 2, 1),
                       _block("conv2.1", 32, 32, "subm2"), _block("conv2.2", 32, 32, "subm2")]),
            ("conv3", [_block("conv3.0", 32, 64, "spconv3", "spconv", 3, 2, 1),
                       _block("conv3.1", 64, 64, "subm3"), _block("conv3.2", 64, 64, "subm3")]),
            ("conv4", [_block("conv4.0", 64, 64, "spconv4", "spconv", 3, 2, (0, 1, 1)),
                       _block("conv4.1", 64, 64, "subm4"), _block("conv4.2", 64, 64, "subm4")]),
        ]
    elif name == "VoxelResBackBone8x":
        c4 = 128
        stages = [
            ("conv_input", [dict(op="conv", conv="conv_input.0", bn="conv_input.1", cin=in_ch, cout=16, key="subm1",
                                 kind="subm", ksize=3, stride=1, pad=1, relu=True)]),
            ("conv1", [_basic("conv1.0", 16, "res1"), _basic("conv1.1", 16, "res1")]),
            ("conv2", [_block("conv2.0", 16, 32, "spconv2", "spconv", 3, 2, 1),
                       _basic("conv2.1", 32, "res2"), _basic("conv2.2", 32, "res2")]),
            ("conv3", [_block("conv3.0", 32, 64, "spconv3", "spconv", 3, 2, 1),
                       _basic("conv3.1", 64, "res3"), _basic("conv3.2", 64, "res3")]),
            ("conv4", [_block("conv4.0", 64, 128, "spconv4", "spconv", 3, 2, (0, 1, 1)),
                       _basic("conv4.1", 128, "res4"), _basic("conv4.2", 128, "res4")]),
        ]
    else:
        raise KeyError(name)
    stages.append(("conv_out", [dict(op="conv", conv="conv_out.0", bn="conv_out.1", cin=c4, cout=128,
                                     key="spconv_down2", kind="spconv", ksize=(3, 1, 1), stride=(2, 1, 1),
                                     pad=last_pad, relu=True)]))
    return stages


class _Sp:
    def __init__(self, features, indices, shape, batch):
        self.features, self.indices, self.shape, self.batch = features, indices, list(shape), batch


def _bn_of(params, prefix):
    return tuple(np.asarray(params[prefix + "." + k], np.float32)
                 for k in ("weight", "bias", "running_mean", "running_var"))


def backbone_forward(name, params, voxel_features, voxel_coords, batch_size, sparse_shape, last_pad=0):
    """Eval-mode forward of the named backbone on CPU.

    params: {state_dict key: ndarray}; voxel_coords [N,4] (b,z,y,x).
    Returns dict with 'x_conv1'..'x_conv4', 'out' -> (features, indices, spatial_shape) and
    'rulebooks' -> {indice_key: (outids, pairs, num)} in the reference's CPU ordering.
    """
    x = _Sp(_f32(voxel_features), _i32(voxel_coords), sparse_shape, batch_size)
    rulebooks = {}

    def conv(x, prefix, kind, key, ksize, stride, pad):
        ks, st, pd = _triple(ksize), _triple(stride), _triple(pad)
        if key in rulebooks:  # conv.py:165-166 cache hit
            outids, pairs, num, oshape = rulebooks[key]
        elif kind == "subm":
            outids, pairs, num = rulebook_subm(x.indices, x.batch, x.shape, ks, 1)
            oshape = list(x.shape)
            rulebooks[key] = (outids, pairs, num, oshape)
        else:
            outids, pairs, num, oshape = rulebook_conv(x.indices, x.batch, x.shape, ks, st, pd, 1)
            rulebooks[key] = (outids, pairs, num, oshape)
        w = params[prefix + ".weight"]
        f = indice_conv(x.features, w, pairs, num, outids.shape[0], False, kind == "subm")
        return _Sp(f, outids, oshape, x.batch)

    outs = {}
    for stage, ops in backbone_plan(name, x.features.shape[1], last_pad):
        for op in ops:
            if op["op"] == "conv":
                y = conv(x, op["conv"], op["kind"], op["key"], op["ksize"], op["stride"], op["pad"])
                y.features = bias_bn_res_relu(y.features, params.get(op["conv"] + ".bias"),
                                              _bn_of(params, op["bn"]), None, op["relu"])
                x = y
            else:  # SparseBasicBlock
                p = op["prefix"]
                identity = x.features
                y = conv(x, p + ".conv1", "subm", op["key"], 3, 1, 1)
                y.features = bias_bn_res_relu(y.features, params.get(p + ".conv1.bias"), _bn_of(params, p + ".bn1"),
                                              None, True)
                z = conv(y, p + ".conv2", "subm", op["key"], 3, 1, 1)
                z.features = bias_bn_res_relu(z.features, params.get(p + ".conv2.bias"), _bn_of(params, p + ".bn2"),
                                              identity, True)
                x = z
        tag = {"conv1": "x_conv1", "conv2": "x_conv2", "conv3": "x_conv3", "conv4": "x_conv4",
               "conv_out": "out"}.get(stage)
        if tag:
            outs[tag] = (x.features, x.indices, list(x.shape))
    outs["rulebooks"] = {k: v[:3] for k, v in rulebooks.items()}
    return outs


# --------------------------------------------------------------------------------------------
# consumers of the sparse outputs: voxel -> point 3-NN interpolation, voxel query
# --------------------------------------------------------------------------------------------
def voxel_centers(voxel_coords_zyx, downsample_times, voxel_size, point_cloud_range):
    """get_voxel_centers (pcdet/utils/common_utils.py:76-92): fp32, (idx + 0.5) * (voxel_size * ds) + range_min."""
    c = np.asarray(voxel_coords_zyx)[:, [2, 1, 0]].astype(np.float32)
    vs = (np.asarray(voxel_size, np.float32) * np.float32(downsample_times)).astype(np.float32)
    lo = np.asarray(point_cloud_range[0:3], np.float32)
    return ((c + np.float32(0.5)) * vs + lo).astype(np.float32)


def three_nn(unknown, known, chunk=2048):
    """three_nn_kernel_fast (pcdet/ops/pointnet2/pointnet2_batch/src/interpolate_gpu.cu:16-58) + the sqrt of
    ThreeNN.forward: for every `unknown` row the three `known` rows with the smallest squared distance, ties by
    lowest index (the kernel scans ascending with strict <).  The differences are rounded to fp32 like the kernel's;
    squares and sum are carried in float64 (the kernel's fp32 FMA chain differs from this by <= 2 ulp, which only
    matters for candidates closer than that to a tie).  Fewer than three known rows: distance inf, index 0.
    Returns (dist [n,3] fp32, idx [n,3] int32, dist2 [n,3] float64)."""
    unknown = _f32(unknown).reshape(-1, 3)
    known = _f32(known).reshape(-1, 3)
    n, m = unknown.shape[0], known.shape[0]
    d2 = np.full((n, 3), np.inf, np.float64)
    idx = np.zeros((n, 3), np.int32)
    if m:
        for s in range(0, n, chunk):
            diff = (unknown[s:s + chunk, None, :] - known[None, :, :]).astype(np.float32).astype(np.float64)
            d = (diff * diff).sum(-1)
            # the 16 smallest (any order), put in ascending index order, then a stable sort by distance: the three
            # smallest by (distance, index), like the kernel's strict-< ascending scan
            kk = min(16, m)
            cand = np.sort(np.argpartition(d, kk - 1, axis=1)[:, :kk], axis=1)
            dc = np.take_along_axis(d, cand, 1)
            pick = np.argsort(dc, axis=1, kind="stable")[:, :3]
            order = np.take_along_axis(cand, pick, 1)
            k = order.shape[1]
            d2[s:s + chunk, :k] = np.take_along_axis(d, order, 1)
            idx[s:s + chunk, :k] = order
    return np.sqrt(d2).astype(np.float32), idx, d2


def top3_interpolate(xyz, new_xyz, feats):
    """top3_interpolate (pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py:292-326): inverse-distance weights
    over the three nearest `xyz` of every `new_xyz`, weighted sum of `feats` (three_interpolate,
    interpolate_gpu.cu:78-100).  Returns (interpolated [M,C] fp32, dist, idx)."""
    dist, idx, _ = three_nn(new_xyz, xyz)
    recip = (np.float32(1.0) / (dist + np.float32(1e-8))).astype(np.float32)
    norm = recip.sum(1, keepdims=True, dtype=np.float32)
    w = (recip / norm).astype(np.float32)
    feats = _f32(feats)
    if feats.shape[0] == 0:
        return np.zeros((dist.shape[0], feats.shape[1]), np.float32), dist, idx
    out = (w[:, 0:1] * feats[idx[:, 0]] + w[:, 1:2] * feats[idx[:, 1]] + w[:, 2:3] * feats[idx[:, 2]])
    return out.astype(np.float32), dist, idx


def voxel_to_point_interpolate(voxel_indices, voxel_feats, point_coords, batch_size, voxel_size, point_cloud_range,
                               downsample_times):
    """Steps 1-3 of ResidualVoxelToPointDecoder.forward (residual_v2p_decoder.py:86-116): per frame, centres of the
    frame's voxels, 3-NN of the frame's points, interpolation.  voxel_indices [N,4] (b,z,y,x), point_coords [P,4]
    (b,x,y,z).  Returns (feats [P,C], dist [P,3], idx [P,3] frame-local)."""
    voxel_indices = _i32(voxel_indices)
    point_coords = _f32(point_coords)
    voxel_feats = _f32(voxel_feats)
    out = np.zeros((point_coords.shape[0], voxel_feats.shape[1]), np.float32)
    dist = np.zeros((point_coords.shape[0], 3), np.float32)
    idx = np.zeros((point_coords.shape[0], 3), np.int32)
    for b in range(int(batch_size)):
        vm = voxel_indices[:, 0] == b
        pm = point_coords[:, 0].astype(np.int64) == b
        xyz = voxel_centers(voxel_indices[vm][:, 1:4], downsample_times, voxel_size, point_cloud_range)
        o, d, i = top3_interpolate(xyz, point_coords[pm][:, 1:4], voxel_feats[vm])
        out[pm], dist[pm], idx[pm] = o, d, i
    return out, dist, idx


def voxel2pinds(indices, spatial_shape, batch_size):
    """generate_voxel2pinds (pcdet/utils/spconv_utils.py:13-21): dense [B,Z,Y,X] int32 grid of row ids, -1 elsewhere."""
    indices = _i32(indices)
    grid = -np.ones([int(batch_size)] + [int(s) for s in spatial_shape], np.int32)
    grid[indices[:, 0], indices[:, 1], indices[:, 2], indices[:, 3]] = np.arange(indices.shape[0], dtype=np.int32)
    return grid


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """voxel_query_kernel_stack (pcdet/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-88) + the empty-ball
    handling of VoxelQuery.forward (voxel_query_utils.py:33-42).  Pure-python loops: small cases only.
    Returns (idx [M,nsample] int32, empty_ball_mask [M] bool)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    new_coords = _i32(new_coords)
    B, R1, R2, R3 = point_indices.shape
    zr, yr, xr = (int(v) for v in max_range)
    M = new_coords.shape[0]
    idx = np.zeros((M, int(nsample)), np.int32)
    r2 = np.float32(radius) * np.float32(radius)
    for pt in range(M):
        b, cz, cy, cx = (int(v) for v in new_coords[pt])
        cnt = 0
        for dz in range(-zr, zr + 1):
            z = cz + dz
            if z < 0 or z >= R1:
                continue
            for dy in range(-yr, yr + 1):
                y = cy + dy
                if y < 0 or y >= R2:
                    continue
                for dx in range(-xr, xr + 1):
                    x = cx + dx
                    if x < 0 or x >= R3:
                        continue
                    nb = int(point_indices[b, z, y, x])
                    if nb < 0:
                        continue
                    diff = (xyz[nb] - new_xyz[pt]).astype(np.float32).astype(np.float64)
                    if float((diff * diff).sum()) > float(r2):
                        continue
                    if cnt < nsample:
                        if cnt == 0:
                            idx[pt, :] = nb
                        idx[pt, cnt] = nb
                        cnt += 1
        if cnt == 0:
            idx[pt, 0] = -1
    empty = idx[:, 0] == -1
    idx[empty] = 0
    return idx, empty
