"""HeightCompression -- same module as the reference's map_to_bev step
(pcdet/models/backbones_2d/map_to_bev/height_compression.py:4-26): the stride-8 sparse tensor densified to
[B, C, D, H, W] and viewed as the BEV feature map [B, C*D, H, W] the 2D backbone reads.

On CUDA the densification is one fv2p_height_compression call (a cell -> row map, then one coalesced pass that
writes every element of the map once); the view is free because [B, C, D, H, W] and [B, C*D, H, W] are the same
bytes.  HotPath(..., bev=True) appends the same call to
the graph-captured step, with the row count read on the device.
"""
import torch
import torch.nn as nn

from . import _lib


def height_compression(features, indices, spatial_shape, batch_size, out=None, n_dev=None, workspace=None):
    """features [N, C] fp32/bf16 CUDA, indices [N, 4] int32 (b, z, y, x) -> [B, C*D, H, W] (zeros elsewhere)."""
    dev = _lib.require_device(features)
    if features.dtype not in (torch.float32, torch.bfloat16) or indices.dtype != torch.int32:
        raise ValueError("height_compression expects fp32/bf16 features and int32 indices")
    features, indices = features.contiguous(), indices.contiguous()
    d, h, w = (int(s) for s in spatial_shape)
    c = features.shape[1]
    if out is None:
        out = torch.empty((int(batch_size), c * d, h, w), dtype=features.dtype, device=features.device)
    if tuple(out.shape) != (int(batch_size), c * d, h, w) or out.dtype != features.dtype or not out.is_contiguous():
        raise ValueError("height_compression: output buffer has the wrong shape, dtype or layout")
    lib = _lib.load()
    shape3 = _lib.i32x3([d, h, w])
    if workspace is None:  # the cell -> row map; callers inside a captured graph pass their own fixed buffer
        workspace = _lib.Workspace.get(features.device, lib.fv2p_height_compression_workspace_bytes(int(batch_size),
                                                                                                      shape3), "bev")
    with torch.cuda.device(dev):
        st = lib.fv2p_height_compression(_lib.ptr(features), _lib.ptr(indices), features.shape[0], _lib.ptr(n_dev),
                                         int(batch_size), c, shape3, features.element_size(), _lib.ptr(out),
                                         _lib.ptr(workspace), workspace.numel(), _lib.stream_ptr(features.device))
    _lib.check(st, "height_compression")
    return out


class _HeightCompressionFn(torch.autograd.Function):
    """Differentiable wrapper: the reference's dense() is a scatter_nd (structure.py:5-18,57-66), which autograd
    follows back to the features; here forward is the CUDA kernel and backward gathers grad[b, :, z, y, x] per row,
    so the 2D backbone's loss still reaches the sparse convolutions in training mode."""

    @staticmethod
    def forward(ctx, features, indices, spatial_shape, batch_size):
        ctx.save_for_backward(indices)
        ctx.shape5 = (int(batch_size), features.shape[1]) + tuple(int(s) for s in spatial_shape)
        return height_compression(features.detach(), indices, spatial_shape, batch_size)

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        ind = indices.long()
        g5 = grad.reshape(ctx.shape5)
        return g5[ind[:, 0], :, ind[:, 1], ind[:, 2], ind[:, 3]].contiguous(), None, None, None


def height_compression_autograd(features, indices, spatial_shape, batch_size):
    """height_compression() that autograd can differentiate with respect to `features`."""
    if torch.is_grad_enabled() and features.requires_grad:
        return _HeightCompressionFn.apply(features, indices, spatial_shape, batch_size)
    return height_compression(features, indices, spatial_shape, batch_size)


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = _cfg_get(model_cfg, 'NUM_BEV_FEATURES')

    def forward(self, batch_dict):
        """batch_dict: encoded_spconv_tensor (+ _stride) -> adds spatial_features [B, C*D, H, W] and
        spatial_features_stride (height_compression.py:10-25)."""
        enc = batch_dict['encoded_spconv_tensor']
        if enc.features.is_cuda and enc.features.dtype in (torch.float32, torch.bfloat16) and \
                enc.indices.dtype == torch.int32 and len(enc.spatial_shape) == 3:
            spatial_features = height_compression_autograd(enc.features, enc.indices, enc.spatial_shape,
                                                           enc.batch_size)
        else:
            dense = enc.dense()
            n, c, d, h, w = dense.shape
            spatial_features = dense.view(n, c * d, h, w)
        batch_dict['spatial_features'] = spatial_features
        batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
        return batch_dict
