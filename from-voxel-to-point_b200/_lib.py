"""ctypes binding of libfv2p_b200.so (include/fv2p_b200.h).  PyTorch is only used for device memory
and streams; every compute call goes through the C ABI with raw pointers.

There is no CPU fallback: a missing shared object, a non-CUDA tensor or a non-sm_100 device raises.
Error mapping mirrors the reference (include/tensorview/tensorview.h:71-102): invalid arguments ->
ValueError, everything else -> RuntimeError.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfv2p_b200.so")

MODE_F32, MODE_BF16_TC, MODE_TF32X3_TC, MODE_BF16_SIMT, MODE_F32_IN_BF16_OUT = 0, 1, 2, 3, 4
MODE_FP32_TC = MODE_TF32X3_TC  # the fp32 tensor-core mode (split bf16 operands today; the older name stays valid)
MAX_KVOL = 32
ABI_VERSION = 2
FLAG_PREFILLED = 1
PREFILL_TABLE, PREFILL_NBR_ALL, PREFILL_NBR_MIRROR, PREFILL_CONV_WS, PREFILL_GROUP_WS = 1, 2, 3, 4, 5
STATUS_OUT_OVERFLOW, STATUS_VOXEL_OVERFLOW = 1, 2

_c_i64, _c_int, _c_sz, _c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); kept in the order of include/fv2p_b200.h
PROTOTYPES = {
    "fv2p_abi_version": (_c_int, []),
    "fv2p_last_error": (ctypes.c_char_p, []),
    "fv2p_device_check": (_c_int, [_c_vp, _c_vp, _c_vp]),
    "fv2p_voxelize_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_i64, _c_int, _c_i64]),
    "fv2p_voxelize_mean": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_int,
                                    _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp, _c_vp]),
    "fv2p_voxelize_mean_table": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_int,
                                    _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "fv2p_voxel_generate": (_c_int, [_c_vp, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp,
                                     _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_mean_vfe": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_int, _c_vp, _c_vp]),
    "fv2p_geometry_prefill": (_c_int, [_c_vp, _c_int, _c_vp]),
    "fv2p_table_bytes": (_c_sz, [_c_i64]),
    "fv2p_table_build": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_int, _c_vp]),
    "fv2p_subm_neighbours": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64,
                                      _c_int, _c_vp]),
    "fv2p_conv_neighbours_workspace_bytes": (_c_sz, [_c_i64, _c_int]),
    "fv2p_conv_neighbours": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                      _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_int, _c_vp]),
    "fv2p_pairs_workspace_bytes": (_c_sz, [_c_i64, _c_int]),
    "fv2p_subm_pairs": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp,
                                 _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_conv_pairs": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp,
                                 _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_rulebook_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "fv2p_rulebook_subm": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp,
                                    _c_i64, _c_vp, _c_sz, _c_vp, _c_vp]),
    "fv2p_rulebook_conv": (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                    _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp, _c_vp]),
    "fv2p_get_indice_pairs_3d": (_c_int, [_c_vp, _c_i64, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                          _c_int, _c_int, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp,
                                          _c_sz, _c_vp]),
    "fv2p_pairs_to_nbr": (_c_int, [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_i64, _c_vp, _c_i64, _c_vp]),
    "fv2p_conv_fwd": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_int, _c_i64, _c_vp, _c_int,
                               _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp]),
    "fv2p_group_rows_workspace_bytes": (_c_sz, [_c_i64]),
    "fv2p_group_rows": (_c_int, [_c_vp, _c_i64, _c_int, _c_int, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz,
                                 _c_int, _c_vp]),
    "fv2p_sort_rows_workspace_bytes": (_c_sz, [_c_i64]),
    "fv2p_sort_rows_by_mask": (_c_int, [_c_vp, _c_i64, _c_int, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz,
                                        _c_vp]),
    "fv2p_tc_gather_mode": (_c_int, [_c_int]),
    "fv2p_pack_weight_bytes": (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    "fv2p_pack_weight": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "fv2p_indice_conv_workspace_bytes": (_c_sz, [_c_int, _c_i64]),
    "fv2p_indice_conv_fp32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int,
                                       _c_int, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_conv_grad_filters": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_i64, _c_vp, _c_int, _c_int, _c_vp, _c_vp]),
    "fv2p_batchnorm_workspace_bytes": (_c_sz, [_c_int]),
    "fv2p_batchnorm_train_fwd": (_c_int, [_c_vp, _c_i64, _c_int, _c_vp, _c_vp, ctypes.c_float, ctypes.c_float, _c_vp,
                                          _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_batchnorm_train_bwd": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                          _c_vp, _c_sz, _c_vp]),
    "fv2p_dense_ncdhw": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_vp]),
    "fv2p_height_compression_workspace_bytes": (_c_sz, [_c_int, _c_vp]),
    "fv2p_height_compression": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_int, _c_int, _c_vp, _c_int, _c_vp, _c_vp, _c_sz,
                                         _c_vp]),
    "fv2p_voxel_three_nn_workspace_bytes": (_c_sz, [_c_int, _c_i64]),
    "fv2p_voxel_three_nn": (_c_int, [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp,
                                     _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_sz, _c_vp]),
    "fv2p_voxel_query": (_c_int, [_c_i64, _c_vp, _c_int, ctypes.c_float, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                  _c_vp, _c_vp]),
    "fv2p_copy_rows": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "fv2p_cast_f32_to_bf16": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp]),
    "fv2p_cast_bf16_to_f32": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp]),
}

_lib = None
_checked_devices = set()


def load():
    """Loads the shared object (no device needed).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libfv2p_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python from-voxel-to-point_b200/build.py`. There is no CPU or PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.fv2p_abi_version() != ABI_VERSION:
            raise RuntimeError("libfv2p_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def last_error():
    return load().fv2p_last_error().decode("utf-8", "replace")


def check(status, what):
    if status == 0:
        return
    msg = "%s failed (%d): %s" % (what, status, last_error())
    if status == -1:
        raise ValueError(msg)
    if status == -3:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def require_device(t):
    """The tensor must live on an sm_100 CUDA device; makes that device current-checked once."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("fv2p_b200 runs on CUDA tensors only (got %s); there is no CPU fallback" %
                         (t.device if isinstance(t, torch.Tensor) else type(t)))
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        with torch.cuda.device(idx):
            check(load().fv2p_device_check(None, None, None), "device check")
        _checked_devices.add(idx)
    return idx


class PrefillItem(ctypes.Structure):
    """fv2p_prefill_item (include/fv2p_b200.h)."""
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("ptr", ctypes.c_void_p),
                ("a", ctypes.c_int64), ("b", ctypes.c_int64)]


def prefill_items(items):
    """[(kind, tensor_or_ptr, a, b), ...] -> (ctypes array, count)."""
    arr = (PrefillItem * max(len(items), 1))()
    for i, (kind, t, a, b) in enumerate(items):
        arr[i].kind, arr[i].reserved = int(kind), 0
        arr[i].ptr = t.data_ptr() if isinstance(t, torch.Tensor) else int(t)
        arr[i].a, arr[i].b = int(a), int(b)
    return arr, len(items)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def i32x3(v):
    if isinstance(v, (list, tuple)) or hasattr(v, "__len__"):
        vals = [int(x) for x in v]
        if len(vals) != 3:
            raise ValueError("expected 3 values, got %r" % (v,))
    else:
        vals = [int(v)] * 3
    return (ctypes.c_int32 * 3)(*vals)


def f32arr(v):
    vals = [float(x) for x in v]
    return (ctypes.c_float * len(vals))(*vals)


class Workspace:
    """Grow-only device scratch buffer per (device, tag); the library never allocates."""
    _pool = {}

    @classmethod
    def get(cls, device, nbytes, tag="default"):
        key = (str(device), tag)
        buf = cls._pool.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            cls._pool[key] = buf
        return buf
