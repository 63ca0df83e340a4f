"""Access to the REAL reference for pinning the oracle and for the reference CPU arm.
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Two levels:

* ``load_ext()``: the reference's compiled ``sparse_conv_ext`` from ``oracle/_ref/`` (built by
  ``oracle/build_ref.py``).  Works wherever the prebuilt file travelled to (this container and the
  GPU box).  ``ext_backbone_forward`` drives the two backbones through that extension with the
  exact call sequence of conv.py:113-229 / spconv_backbone.py:134-186,241-290 (rulebook cache per
  indice_key, indice_conv_fp32, bias, BatchNorm1d eval, residual, ReLU) using torch CPU ops.
* ``load_reference_python()``: additionally imports the reference's *Python* files in place from
  /root/reference (numba voxelizer, MeanVFE, spconv package, backbones).  Only possible in the build
  container; used by tests/golden/make_golden.py to generate the committed fixtures.
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

from . import build_ref
from .oracle import backbone_plan, conv_output_size, _triple

REF_ROOT = build_ref.REF_ROOT
_ext = None


def have_ext():
    return build_ref.built_path() is not None


def load_ext():
    """Import oracle/_ref/sparse_conv_ext*.so (pybind module of src/all.cc:22-71)."""
    global _ext
    if _ext is None:
        path = build_ref.built_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref/sparse_conv_ext.so missing: run `python oracle/build_ref.py` "
                                    "in the build container")
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        loader = importlib.machinery.ExtensionFileLoader(build_ref.EXT_NAME, path)
        spec = importlib.util.spec_from_loader(build_ref.EXT_NAME, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _ext = mod
    return _ext


# --------------------------------------------------------------------------------------------
# reference extension driven directly (runs on the GPU box too)
# --------------------------------------------------------------------------------------------
def ext_get_indice_pairs(indices, batch, in_shape, ksize, stride, pad, dil, subm):
    """ops.py:46-94 -> sparse_conv_ext.get_indice_pairs_3d (spconv_ops.h:28)."""
    ext = load_ext()
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(pad), _triple(dil)
    in_shape = [int(s) for s in in_shape]
    out_shape = in_shape if subm else conv_output_size(in_shape, ks, st, pd, dl)
    outids, pairs, num = ext.get_indice_pairs_3d(indices, int(batch), out_shape, in_shape, ks, st, pd, dl,
                                                 [0, 0, 0], int(subm), 0)
    return outids, pairs, num, out_shape


def ext_backbone_forward(name, params, voxel_features, voxel_coords, batch_size, sparse_shape, last_pad=0,
                         timers=None, device=None):
    """Eval forward of VoxelBackBone8x / VoxelResBackBone8x through the reference extension.

    params: {state_dict key: torch.Tensor (fp32)}.  Returns the same structure as oracle.backbone_forward but
    with torch tensors.  device=None: CPU tensors, the reference's CPU functors (the parity target).
    device='cuda': the same call sequence on CUDA tensors, i.e. the reference's own GPU kernels
    (src/indice_cuda.cu:30-135, src/reordering_cuda.cu:31-140, cuBLAS through torch::mm_out) - used by
    bench.py's `reference_gpu` block only; its rulebook ordering differs from the CPU path.
    """
    import time

    import torch
    import torch.nn.functional as F

    ext = load_ext()
    feats = torch.as_tensor(voxel_features, dtype=torch.float32)
    inds = torch.as_tensor(voxel_coords).int().contiguous()
    if device is not None:
        feats, inds = feats.to(device), inds.to(device)
    shape = [int(s) for s in sparse_shape]
    books = {}
    t_rb = t_cv = 0.0

    def conv(feats, inds, shape, prefix, kind, key, ksize, stride, pad):
        nonlocal t_rb, t_cv
        t0 = time.perf_counter()
        if key in books:
            outids, pairs, num, oshape = books[key]
        else:
            outids, pairs, num, oshape = ext_get_indice_pairs(inds, batch_size, shape, ksize, stride, pad, 1,
                                                              kind == "subm")
            books[key] = (outids, pairs, num, oshape)
        t1 = time.perf_counter()
        w = params[prefix + ".weight"]
        out = ext.indice_conv_fp32(feats, w, pairs, num, outids.shape[0], 0, int(kind == "subm"))
        b = params.get(prefix + ".bias")
        if b is not None:
            out += b
        t_rb += t1 - t0
        t_cv += time.perf_counter() - t1
        return out, outids, oshape

    def bn(x, prefix):
        return F.batch_norm(x, params[prefix + ".running_mean"], params[prefix + ".running_var"],
                            params[prefix + ".weight"], params[prefix + ".bias"], False, 0.01, 1e-3)

    outs = {}
    with torch.no_grad():
        for stage, ops in backbone_plan(name, feats.shape[1], last_pad):
            for op in ops:
                if op["op"] == "conv":
                    feats, inds, shape = conv(feats, inds, shape, op["conv"], op["kind"], op["key"], op["ksize"],
                                              op["stride"], op["pad"])
                    feats = torch.relu(bn(feats, op["bn"]))
                else:
                    p = op["prefix"]
                    identity = feats
                    f1, inds, shape = conv(feats, inds, shape, p + ".conv1", "subm", op["key"], 3, 1, 1)
                    f1 = torch.relu(bn(f1, p + ".bn1"))
                    f2, inds, shape = conv(f1, inds, shape, p + ".conv2", "subm", op["key"], 3, 1, 1)
                    f2 = bn(f2, p + ".bn2")
                    f2 += identity
                    feats = torch.relu(f2)
            tag = {"conv1": "x_conv1", "conv2": "x_conv2", "conv3": "x_conv3", "conv4": "x_conv4",
                   "conv_out": "out"}.get(stage)
            if tag:
                outs[tag] = (feats, inds, list(shape))
    outs["rulebooks"] = {k: v[:3] for k, v in books.items()}
    if timers is not None:
        timers["rulebook_s"] = timers.get("rulebook_s", 0.0) + t_rb
        timers["conv_s"] = timers.get("conv_s", 0.0) + t_cv
    return outs


# --------------------------------------------------------------------------------------------
# reference Python imported in place (build container only)
# --------------------------------------------------------------------------------------------
def have_reference_python():
    return os.path.isdir(os.path.join(REF_ROOT, "pcdet", "ops", "spconv"))


def _stub_pkg(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load_reference_python():
    """Returns a namespace with the reference's VoxelGenerator, MeanVFE, spconv package and backbones.

    Parent packages are registered as bare stubs whose __path__ points into /root/reference so that the
    reference files are imported unmodified without executing pcdet/__init__.py (which needs easydict,
    tensorboardX, skimage -- absent here).  mmcv.cnn.CONV_LAYERS (conv.py:17) is stubbed: it is a
    registry decorator with no arithmetic.
    """
    if not have_reference_python():
        raise FileNotFoundError(REF_ROOT + " is not present (only exists in the build container)")
    if "pcdet.ops.spconv" in sys.modules and hasattr(sys.modules["pcdet.ops.spconv"], "SparseConvTensor"):
        return _namespace()
    ext = load_ext()
    cnn = types.ModuleType("mmcv.cnn")

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    cnn.CONV_LAYERS = _Registry()
    mm = types.ModuleType("mmcv")
    mm.cnn = cnn
    sys.modules.setdefault("mmcv", mm)
    sys.modules.setdefault("mmcv.cnn", cnn)
    root = os.path.join(REF_ROOT, "pcdet")
    _stub_pkg("pcdet", root)
    _stub_pkg("pcdet.ops", os.path.join(root, "ops"))
    _stub_pkg("pcdet.models", os.path.join(root, "models"))
    _stub_pkg("pcdet.models.backbones_3d", os.path.join(root, "models", "backbones_3d"))
    _stub_pkg("pcdet.models.backbones_3d.vfe", os.path.join(root, "models", "backbones_3d", "vfe"))
    _stub_pkg("pcdet.datasets", os.path.join(root, "datasets"))
    _stub_pkg("pcdet.datasets.processor", os.path.join(root, "datasets", "processor"))
    sys.modules["pcdet.ops.spconv.sparse_conv_ext"] = ext  # ops.py:17 `from . import sparse_conv_ext`
    sp = importlib.import_module("pcdet.ops.spconv")
    sys.modules["pcdet.ops"].spconv = sp
    return _namespace()


def _namespace():
    ns = types.SimpleNamespace()
    ns.spconv = importlib.import_module("pcdet.ops.spconv")
    ns.VoxelGenerator = importlib.import_module("pcdet.datasets.processor.voxel_generator").VoxelGenerator
    ns.MeanVFE = importlib.import_module("pcdet.models.backbones_3d.vfe.mean_vfe").MeanVFE
    bb = importlib.import_module("pcdet.models.backbones_3d.spconv_backbone")
    ns.VoxelBackBone8x = bb.VoxelBackBone8x
    ns.VoxelResBackBone8x = bb.VoxelResBackBone8x
    ns.ext = load_ext()
    return ns


def to_numpy_params(state_dict):
    return {k: v.detach().cpu().numpy() for k, v in state_dict.items() if v.dtype.is_floating_point}


# --------------------------------------------------------------------------------------------
# reference pointnet2 CUDA kernels (oracle/_ref/libref_pointnet2.so): GPU box only
# --------------------------------------------------------------------------------------------
_pn2 = None


def have_pointnet2():
    return build_ref.pointnet2_path() is not None


def load_pointnet2():
    """ctypes handle on the reference's own kernel launchers (C++ mangled names; built by
    oracle/build_ref.py:build_pointnet2 from pointnet2_batch/src/interpolate_gpu.cu and
    pointnet2_stack/src/voxel_query_gpu.cu).  They launch on the legacy default stream."""
    global _pn2
    if _pn2 is None:
        import ctypes
        path = build_ref.pointnet2_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref_pointnet2.so missing: run `python oracle/build_ref.py`")
        lib = ctypes.CDLL(path)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.three_nn = getattr(lib, "_Z29three_nn_kernel_launcher_fastiiiPKfS0_PfPi")
        lib.three_nn.argtypes = [ci, ci, ci, vp, vp, vp, vp]
        lib.three_nn.restype = None
        lib.three_interpolate = getattr(lib, "_Z38three_interpolate_kernel_launcher_fastiiiiPKfPKiS0_Pf")
        lib.three_interpolate.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp]
        lib.three_interpolate.restype = None
        lib.voxel_query = getattr(lib, "_Z33voxel_query_kernel_launcher_stackiiiiifiiiPKfS0_PKiS2_Pi")
        lib.voxel_query.argtypes = [ci, ci, ci, ci, ci, cf, ci, ci, ci, vp, vp, vp, vp, vp]
        lib.voxel_query.restype = None
        _pn2 = lib
    return _pn2


def ref_three_nn(unknown, known):
    """ThreeNN.forward (pointnet2_batch/pointnet2_utils.py:107-129) on CUDA tensors [n,3], [m,3] (batch of one, as
    top3_interpolate calls it).  Returns (dist [n,3] = sqrt(dist2), idx [n,3] int32)."""
    import torch
    lib = load_pointnet2()
    unknown, known = unknown.contiguous(), known.contiguous()
    n, m = unknown.shape[0], known.shape[0]
    dist2 = torch.empty((1, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((1, n, 3), dtype=torch.int32, device=unknown.device)
    torch.cuda.synchronize()
    lib.three_nn(1, n, m, unknown.data_ptr(), known.data_ptr(), dist2.data_ptr(), idx.data_ptr())
    torch.cuda.synchronize()
    return torch.sqrt(dist2)[0], idx[0]


def ref_top3_interpolate(xyz, new_xyz, feats):
    """top3_interpolate (pointnet2_utils.py:292-326) with the reference's kernels: returns [M, Cf]."""
    import torch
    lib = load_pointnet2()
    dist, idx = ref_three_nn(new_xyz, xyz)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=1, keepdim=True)
    weight = (dist_recip / norm).contiguous()
    feats_b = feats.t().contiguous()  # (Cf, N)
    c, m = feats_b.shape
    n = new_xyz.shape[0]
    out = torch.empty((c, n), dtype=torch.float32, device=feats.device)
    torch.cuda.synchronize()
    lib.three_interpolate(1, c, m, n, feats_b.data_ptr(), idx.contiguous().data_ptr(), weight.data_ptr(),
                          out.data_ptr())
    torch.cuda.synchronize()
    return out.t().contiguous(), dist, idx


def ref_voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """VoxelQuery.forward (pointnet2_stack/voxel_query_utils.py:12-42) with the reference's kernel and its dense
    [B,Z,Y,X] int32 grid."""
    import torch
    lib = load_pointnet2()
    m = new_coords.shape[0]
    _, z, y, x = point_indices.shape
    idx = torch.zeros((m, int(nsample)), dtype=torch.int32, device=xyz.device)
    zr, yr, xr = max_range
    torch.cuda.synchronize()
    lib.voxel_query(m, z, y, x, int(nsample), float(radius), int(zr), int(yr), int(xr),
                    new_xyz.contiguous().data_ptr(), xyz.contiguous().data_ptr(), new_coords.contiguous().data_ptr(),
                    point_indices.contiguous().data_ptr(), idx.data_ptr())
    torch.cuda.synchronize()
    empty = idx[:, 0] == -1
    idx[empty] = 0
    return idx, empty
