"""MeanVFE with the reference's constructor and batch_dict contract
(pcdet/models/backbones_3d/vfe/mean_vfe.py:7-31)."""
import torch

from . import _lib
from .vfe_template import VFETemplate


class MeanVFE(VFETemplate):
    def __init__(self, model_cfg, num_point_features, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, batch_dict, **kwargs):
        """batch_dict['voxels'] [M,T,C], ['voxel_num_points'] [M] -> ['voxel_features'] [M,C].

        When the batch was produced by this package's fused voxelizer (BatchVoxelizer / HotPath) the mean
        is already in batch_dict['voxel_features'] and no padded 'voxels' tensor exists: pass through.
        """
        if 'voxels' not in batch_dict and 'voxel_features' in batch_dict:
            return batch_dict
        voxels, num = batch_dict['voxels'], batch_dict['voxel_num_points']
        dev = _lib.require_device(voxels)
        if voxels.dtype != torch.float32:
            raise ValueError("MeanVFE: voxels must be float32")
        voxels = voxels.contiguous()
        num = num.to(torch.int32).contiguous()  # load_data_to_gpu turned the counts into floats (models/__init__.py:15-21)
        m, t, c = voxels.shape
        out = torch.empty((m, c), dtype=torch.float32, device=voxels.device)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().fv2p_mean_vfe(_lib.ptr(voxels), _lib.ptr(num), m, t, c, _lib.ptr(out),
                                                 _lib.stream_ptr(voxels.device)), "mean_vfe")
        batch_dict['voxel_features'] = out
        return batch_dict
