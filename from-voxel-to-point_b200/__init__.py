"""fv2p_b200 -- B200-native voxelization + sparse 3D convolution backbone, a drop-in for the hot path of
jialeli1/From-Voxel-to-Point (VoxelGenerator + MeanVFE + VoxelBackBone8x / VoxelResBackBone8x over
spconv.SparseConvTensor / SparseSequential / SubMConv3d / SparseConv3d).

Import name: ``fv2p_b200`` (the directory is called ``from-voxel-to-point_b200``; ``fv2p_b200.py`` at the
repository root registers it under an importable name).  All compute goes through ``libfv2p_b200.so``
(include/fv2p_b200.h); nothing here falls back to CPU or to plain PyTorch.
"""
from . import _lib, spconv, synth
from .height_compression import HeightCompression
from .mean_vfe import MeanVFE
from .spconv_backbone import SparseBasicBlock, VoxelBackBone8x, VoxelResBackBone8x, post_act_block
from .voxel_generator import BatchVoxelizer, VoxelGenerator
from .engine import BackboneEngine
from .pipeline import HotPath
from . import sharding
from . import pointops
from .data_processor import DataProcessor
from .batchnorm import BatchNorm1d

# name lookup tables like pcdet/models/backbones_3d/__init__.py:6-12 and vfe/__init__.py:5-9
BACKBONES_3D = {'VoxelBackBone8x': VoxelBackBone8x, 'VoxelResBackBone8x': VoxelResBackBone8x}
VFE = {'MeanVFE': MeanVFE}
MAP_TO_BEV = {'HeightCompression': HeightCompression}  # backbones_2d/map_to_bev/__init__.py

__all__ = ['spconv', 'synth', 'MeanVFE', 'VoxelGenerator', 'BatchVoxelizer', 'VoxelBackBone8x',
           'VoxelResBackBone8x', 'SparseBasicBlock', 'post_act_block', 'BackboneEngine', 'HotPath', 'BACKBONES_3D',
           'VFE', 'HeightCompression', 'MAP_TO_BEV', 'pointops', 'DataProcessor', 'BatchNorm1d']
