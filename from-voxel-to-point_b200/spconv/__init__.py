"""Drop-in for ``pcdet.ops.spconv`` on the hot path (reference: pcdet/ops/spconv/__init__.py:15-40).

Exports the names the FV2P / MGAF-3DSSD backbones import -- SparseConvTensor, SparseSequential, SparseModule,
SubMConv3d, SparseConv3d -- plus scatter_nd, ToDense and RemoveGrid.  2D/4D, transposed, inverse, max-pool and
group variants are outside SURVEY.md section 8 and raise NotImplementedError on use.
"""
from . import functional, ops
from .conv import SparseConv3d, SparseConvolution, SubMConv3d
from .modules import RemoveGrid, SparseModule, SparseSequential, ToDense
from .structure import SparseConvTensor, scatter_nd


def _outside_hot_path(name):
    class _Missing(SparseModule):
        def __init__(self, *a, **k):
            raise NotImplementedError(name + " is outside the hot path rebuilt by fv2p_b200 (SURVEY.md section 8)")
    _Missing.__name__ = name
    return _Missing


SparseConv2d = _outside_hot_path("SparseConv2d")
SubMConv2d = _outside_hot_path("SubMConv2d")
SparseConvTranspose2d = _outside_hot_path("SparseConvTranspose2d")
SparseConvTranspose3d = _outside_hot_path("SparseConvTranspose3d")
SparseInverseConv2d = _outside_hot_path("SparseInverseConv2d")
SparseInverseConv3d = _outside_hot_path("SparseInverseConv3d")
SparseMaxPool2d = _outside_hot_path("SparseMaxPool2d")
SparseMaxPool3d = _outside_hot_path("SparseMaxPool3d")
SparseGroup3d = _outside_hot_path("SparseGroup3d")
SubMGroup3d = _outside_hot_path("SubMGroup3d")

__all__ = [
    'SparseConv2d', 'SparseConv3d', 'SubMConv2d', 'SubMConv3d', 'SparseConvTranspose2d', 'SparseConvTranspose3d',
    'SparseInverseConv2d', 'SparseInverseConv3d', 'SparseModule', 'SparseSequential', 'SparseMaxPool2d',
    'SparseMaxPool3d', 'SparseConvTensor', 'scatter_nd', 'SparseGroup3d', 'SubMGroup3d', 'ToDense', 'RemoveGrid',
    'SparseConvolution',
]
