"""Op layer: same entry points as pcdet/ops/spconv/ops.py:20-158, backed by libfv2p_b200.so through
ctypes instead of the pybind ``sparse_conv_ext``."""
import numpy as np
import torch

from .. import _lib


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """ops.py:20-30."""
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        size = (input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
        if kernel_size[i] == -1:
            output_size.append(1)
        else:
            output_size.append(size)
    return output_size


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation, output_padding):
    """ops.py:33-43."""
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        if kernel_size[i] == -1:
            raise ValueError("deconv don't support kernel_size < 0")
        size = (input_size[i] - 1) * stride[i] - 2 * padding[i] + kernel_size[i] + output_padding[i]
        output_size.append(size)
    return output_size


def _listify(v, ndim):
    return list(v) if isinstance(v, (list, tuple)) else [v] * ndim


def candidate_fanout(ksize, stride, padding, dilation):
    """Upper bound of outputs one input can touch (product of per-axis candidate counts)."""
    e = 1
    for k, s, p, d in zip(ksize, stride, padding, dilation):
        best = 0
        for pos in range(4 * s * k * d + p + 8):
            lo = int((pos - (k - 1) * d - 1 + s + p) / s)  # C division truncates toward zero
            hi = (pos + p) // s
            best = max(best, int((hi - lo) / d) + 1)
        e *= best
    return e


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1, out_padding=0,
                     subm=False, transpose=False, grid=None, return_nbr=False, want_pairs=True):
    """ops.py:46-105.  Returns (outids, indice_pairs [K,2,N] int32, indice_pair_num [K] int32), bit-identical
    to the reference's CPU path.  With return_nbr=True also returns the output-major neighbour map [K,Nout];
    want_pairs=False skips the reference-layout pair lists (both come back as None).  `indices` must hold each
    coordinate once (include/fv2p_b200.h, "Coordinates must be UNIQUE")."""
    ndim = indices.shape[1] - 1
    ksize, stride, padding, dilation, out_padding = (_listify(v, ndim) for v in
                                                     (ksize, stride, padding, dilation, out_padding))
    for d, s in zip(dilation, stride):
        assert any([s == 1, d == 1]), "don't support this."
    if ndim != 3:
        raise NotImplementedError("fv2p_b200 builds the 3D rulebooks of the hot path only (ndim=%d)" % ndim)
    if transpose:
        raise NotImplementedError("transposed sparse convolution is outside the hot path")
    if indices.dtype != torch.int32:
        raise ValueError("indices must be int32 (reference: check_torch_dtype, torch_utils.h:29-60)")
    dev = _lib.require_device(indices)
    lib = _lib.load()
    indices = indices.contiguous()
    spatial_shape = [int(s) for s in spatial_shape]
    n = indices.shape[0]
    kvol = int(np.prod(ksize))
    if subm:
        out_shape = spatial_shape
        out_cap = n
    else:
        out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
        out_cap = max(1, min(n * candidate_fanout(ksize, stride, padding, dilation),
                             int(batch_size) * int(np.prod(out_shape))))
    device = indices.device
    pairs = torch.empty((kvol, 2, n), dtype=torch.int32, device=device) if want_pairs else None
    pair_num = torch.empty((kvol,), dtype=torch.int32, device=device) if want_pairs else None
    nbr = torch.empty((kvol, out_cap), dtype=torch.int32, device=device)
    outids = indices if subm else torch.empty((out_cap, 4), dtype=torch.int32, device=device)
    ws_bytes = lib.fv2p_rulebook_workspace_bytes(n, out_cap, kvol) + 256
    ws = _lib.Workspace.get(device, ws_bytes, "rulebook")
    n_out = (_lib.ctypes.c_int32 * 1)(0)
    with torch.cuda.device(dev):
        st = lib.fv2p_get_indice_pairs_3d(_lib.ptr(indices), n, int(batch_size), _lib.i32x3(out_shape),
                                          _lib.i32x3(spatial_shape), _lib.i32x3(ksize), _lib.i32x3(stride),
                                          _lib.i32x3(padding), _lib.i32x3(dilation), _lib.i32x3(out_padding),
                                          int(subm), int(transpose), _lib.ptr(outids), out_cap, _lib.ptr(pairs),
                                          _lib.ptr(pair_num), _lib.ptr(nbr), out_cap, n_out, _lib.ptr(ws),
                                          ws.numel(), _lib.stream_ptr(device))
    _lib.check(st, "get_indice_pairs")
    n_out = int(n_out[0])
    if not subm:
        outids = outids[:n_out]
        nbr = nbr[:, :n_out]
    if return_nbr:
        return outids, pairs, pair_num, nbr
    return outids, pairs, pair_num


def pairs_to_nbr(indice_pairs, indice_pair_num, num_activate_out, inverse=False):
    """Output-major neighbour map [K,Nout] from a reference-layout pair tensor."""
    dev = _lib.require_device(indice_pairs)
    pairs = indice_pairs.contiguous()
    num = indice_pair_num.to(pairs.device).contiguous()
    kvol, _, stride = pairs.shape
    nbr = torch.empty((kvol, max(int(num_activate_out), 1)), dtype=torch.int32, device=pairs.device)
    with torch.cuda.device(dev):
        st = _lib.load().fv2p_pairs_to_nbr(_lib.ptr(pairs), _lib.ptr(num), kvol, stride, int(inverse),
                                           int(num_activate_out), _lib.ptr(nbr), nbr.shape[1],
                                           _lib.stream_ptr(pairs.device))
    _lib.check(st, "pairs_to_nbr")
    return nbr[:, :int(num_activate_out)]


def pack_weight(weight_flat, mode):
    """Packed shared-memory image of W [K,Cin,Cout] for the tensor-core modes (fv2p_pack_weight)."""
    dev = _lib.require_device(weight_flat)
    lib = _lib.load()
    kvol, cin, cout = weight_flat.shape
    nbytes = lib.fv2p_pack_weight_bytes(kvol, cin, cout, int(mode))
    if nbytes == 0:
        raise ValueError("tensor-core conv needs cin in {16,32,64,128,256} and cout in {16,32,64,128}")
    w = weight_flat.detach().float().contiguous()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    with torch.cuda.device(dev):
        _lib.check(lib.fv2p_pack_weight(_lib.ptr(w), kvol, cin, cout, int(mode), _lib.ptr(packed),
                                        _lib.stream_ptr(w.device)), "pack_weight")
    packed.fv2p_cout = cout
    return packed


def sort_rows_by_mask(nbr, num_activate_out=None, return_tile_order=False):
    """fv2p_sort_rows_by_mask: (perm [Nout], nbr_sorted [K,Nout]) for the tensor-core conv modes, plus the
    128-row tiles as (tile, offset mask) pairs by descending number of active offsets (``tile_order`` [T,2]) if
    asked for."""
    dev = _lib.require_device(nbr)
    kvol = nbr.shape[0]
    n = int(nbr.shape[1] if num_activate_out is None else num_activate_out)
    assert nbr.stride(1) == 1
    perm = torch.empty((max(n, 1),), dtype=torch.int32, device=nbr.device)
    nbr_sorted = torch.empty((kvol, max(n, 1)), dtype=torch.int32, device=nbr.device)
    order = torch.empty(((n + 127) // 128 + 1, 2), dtype=torch.int32, device=nbr.device) if return_tile_order else None
    lib = _lib.load()
    ws = _lib.Workspace.get(nbr.device, lib.fv2p_sort_rows_workspace_bytes(n), "sort")
    with torch.cuda.device(dev):
        st = lib.fv2p_sort_rows_by_mask(_lib.ptr(nbr), nbr.stride(0), kvol, n, None, _lib.ptr(perm),
                                        _lib.ptr(nbr_sorted), nbr_sorted.stride(0), _lib.ptr(order), _lib.ptr(ws),
                                        ws.numel(), _lib.stream_ptr(nbr.device))
    _lib.check(st, "sort_rows_by_mask")
    if return_tile_order:
        return perm[:n], nbr_sorted[:, :n], order[:(n + 127) // 128]
    return perm[:n], nbr_sorted[:, :n]


def conv_forward(features, weight_flat, nbr, num_activate_out, bias=None, scale=None, shift=None, residual=None,
                 relu=False, mode=None, n_out_dev=None, out=None, row_perm=None, tile_order=None, dynamic=True):
    """fv2p_conv_fwd: out = act((sum_k X[nbr[k]] W[k] + bias) * scale + shift + residual).

    ``row_perm`` / ``tile_order`` come from sort_rows_by_mask (tensor-core modes); ``dynamic`` hands the tiles to the
    CTAs from an atomic counter (two zeroed scheduler words per call) instead of round-robin."""
    dev = _lib.require_device(features)
    if mode is None:
        mode = _lib.MODE_F32 if features.dtype == torch.float32 else _lib.MODE_BF16_SIMT
    kvol = nbr.shape[0]
    cin = features.shape[1]
    if mode in (_lib.MODE_F32, _lib.MODE_BF16_SIMT, _lib.MODE_F32_IN_BF16_OUT):
        assert weight_flat.dtype == torch.float32 and weight_flat.shape[0] == kvol and weight_flat.shape[1] == cin
        cout = weight_flat.shape[2]
    else:
        cout = int(weight_flat.fv2p_cout)
    out_dtype = torch.float32 if mode in (_lib.MODE_F32, _lib.MODE_TF32X3_TC) else torch.bfloat16
    n_cap = int(num_activate_out)
    if out is None:
        out = torch.empty((n_cap, cout), dtype=out_dtype, device=features.device)
    assert nbr.stride(1) == 1 and out.is_contiguous() and features.is_contiguous()
    tc = mode in (_lib.MODE_BF16_TC, _lib.MODE_TF32X3_TC)
    sched = torch.zeros((2,), dtype=torch.int32, device=features.device) if (tc and dynamic) else None
    with torch.cuda.device(dev):
        st = _lib.load().fv2p_conv_fwd(_lib.ptr(features), features.shape[0], _lib.ptr(weight_flat), _lib.ptr(nbr),
                                       nbr.stride(0), _lib.ptr(row_perm), _lib.ptr(tile_order), _lib.ptr(sched),
                                       kvol, n_cap, _lib.ptr(n_out_dev), cin, cout, _lib.ptr(bias), _lib.ptr(scale),
                                       _lib.ptr(shift), _lib.ptr(residual), int(relu), int(mode), _lib.ptr(out),
                                       _lib.stream_ptr(features.device))
    _lib.check(st, "conv_fwd")
    return out


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse=False, subm=False,
                nbr=None):
    """ops.py:108-126 (indice_conv_fp32 / _half).  fp32 -> fp32 FMA path; bf16 supported on top of the
    reference's dtypes; fp16 is not built."""
    if filters.dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError("indice_conv: dtype %s is not built (fp32 and bf16 are)" % filters.dtype)
    _lib.require_device(features)
    cin, cout = filters.shape[-2], filters.shape[-1]
    w = filters.reshape(-1, cin, cout).float().contiguous()
    if nbr is None:
        nbr = pairs_to_nbr(indice_pairs, indice_pair_num, num_activate_out, inverse)
    feats = features.contiguous()
    mode = _lib.MODE_F32 if feats.dtype == torch.float32 else _lib.MODE_BF16_SIMT
    return conv_forward(feats, w, nbr, num_activate_out, mode=mode)


def fused_indice_conv(features, filters, bias, indice_pairs, indice_pair_num, num_activate_out, inverse, subm):
    """ops.py:129-140: the reference's "fused" op only pre-loads the bias (fused_spconv_ops.h:29-32)."""
    w = filters.reshape(-1, filters.shape[-2], filters.shape[-1]).float().contiguous()
    nbr = pairs_to_nbr(indice_pairs, indice_pair_num, num_activate_out, inverse)
    return conv_forward(features.contiguous(), w, nbr, num_activate_out, bias=bias.float().contiguous())


def indice_conv_backward(features, filters, out_bp, indice_pairs, indice_pair_num, inverse=False, subm=False):
    """ops.py:143-158 / spconv_ops.h:365-457.  NOT on the hot path.

    grad_input is the same contraction as the forward pass with the roles of the two row sets swapped and W
    transposed -- dX[j] = sum_k dY[out(k, j)] W[k]^T -- so it runs through fv2p_conv_fwd on the input-major neighbour
    map (fv2p_pairs_to_nbr with the pair columns swapped); every input row accumulates its offsets in ascending k, no
    atomics.  grad_filters[k] = X[in]^T dY[out] is ONE fv2p_conv_grad_filters launch on the forward map (the reference
    runs a gather + gather + torch::mm per offset inside a host loop over the D2H-copied pair counts,
    spconv_ops.h:378, 399-436).  Nothing here synchronises with the host."""
    cin, cout = filters.shape[-2], filters.shape[-1]
    if not (features.is_cuda and features.dtype == torch.float32 and out_bp.dtype == torch.float32 and
            indice_pairs.dtype == torch.int32):
        raise ValueError("indice_conv_backward runs on CUDA float32 tensors and int32 pairs only (no CPU fallback)")
    dev = _lib.require_device(features)
    w = filters.reshape(-1, cin, cout).float()
    kvol = w.shape[0]
    feats, go = features.contiguous(), out_bp.contiguous()
    n_in, n_out = feats.shape[0], go.shape[0]
    nbr_in = pairs_to_nbr(indice_pairs, indice_pair_num, n_in, inverse=not inverse)
    grad_in = conv_forward(go, w.transpose(1, 2).contiguous(), nbr_in.contiguous(), n_in, mode=_lib.MODE_F32)
    nbr_out = pairs_to_nbr(indice_pairs, indice_pair_num, n_out, inverse=inverse).contiguous()
    grad_w = torch.empty((kvol, cin, cout), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(dev):
        st = _lib.load().fv2p_conv_grad_filters(_lib.ptr(feats), _lib.ptr(go), _lib.ptr(nbr_out), nbr_out.stride(0),
                                                kvol, n_out, None, cin, cout, _lib.ptr(grad_w),
                                                _lib.stream_ptr(feats.device))
    _lib.check(st, "conv_grad_filters")
    return grad_in, grad_w.view_as(filters).to(filters.dtype)
