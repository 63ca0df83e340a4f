"""SparseConvTensor -- same holder as the reference (pcdet/ops/spconv/structure.py:21-71)."""
import numpy as np
import torch

from .. import _lib


def scatter_nd(indices, updates, shape):
    """Dense tensor of ``shape`` with ``updates`` written at ``indices`` (reference helper, structure.py:5-18;
    kept for callers that import it)."""
    lead = indices.shape[-1]
    coords = indices.reshape(-1, lead).long()
    dense = updates.new_zeros(tuple(shape))
    dense[tuple(coords.unbind(1))] = updates.reshape(coords.shape[0], *shape[lead:])
    return dense


class _DenseFn(torch.autograd.Function):
    """dense() through fv2p_dense_ncdhw, differentiable with respect to the features like the reference's
    scatter_nd path: backward gathers grad[b, :, z, y, x] for every row."""

    @staticmethod
    def forward(ctx, features, indices, shape, batch_size):
        ctx.save_for_backward(indices)
        f = features.detach().contiguous()
        out = torch.zeros([batch_size, f.shape[1]] + list(shape), dtype=f.dtype, device=f.device)
        if f.shape[0]:
            _lib.check(_lib.load().fv2p_dense_ncdhw(_lib.ptr(f), _lib.ptr(indices), f.shape[0], None, f.shape[1],
                                                    _lib.i32x3(shape), _lib.ptr(out), _lib.stream_ptr(f.device)),
                       "dense")
        return out

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        ind = indices.long()
        return grad[ind[:, 0], :, ind[:, 1], ind[:, 2], ind[:, 3]].contiguous(), None, None, None


class SparseConvTensor(object):
    """features [N,C], indices [N,ndim+1] int32 (batch first), spatial_shape, batch_size.

    ``indice_dict`` maps indice_key -> (outids, indices, indice_pairs [K,2,N], indice_pair_num [K],
    spatial_shape) exactly like conv.py:180-183 and is shared by reference with every tensor derived from
    this one (conv.py:227).  ``nbr_dict`` is this package's side cache (indice_key -> output-major
    neighbour map consumed by the CUDA conv); it is shared the same way.
    """

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        if self.indices.dtype != torch.int32:
            self.indices.int()  # no-op kept from the reference (structure.py:37-38): callers pass int32
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.nbr_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_pair(self, key):
        if key is None:
            return None
        if key in self.indice_dict:
            return self.indice_dict[key]
        return None

    def dense(self, channels_first=True):
        """structure.py:57-66.  fp32 CUDA features go through fv2p_dense_ncdhw."""
        shape = [int(s) for s in self.spatial_shape]
        f = self.features
        if channels_first and len(shape) == 3 and f.is_cuda and f.dtype == torch.float32 and \
                self.indices.dtype == torch.int32:
            _lib.require_device(f)
            return _DenseFn.apply(f, self.indices.contiguous(), shape, int(self.batch_size))
        output_shape = [self.batch_size] + shape + [f.shape[1]]
        res = scatter_nd(self.indices.long(), f, output_shape)
        if not channels_first:
            return res
        ndim = len(shape)
        trans_params = list(range(0, ndim + 1))
        trans_params.insert(1, ndim + 1)
        return res.permute(*trans_params).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size
