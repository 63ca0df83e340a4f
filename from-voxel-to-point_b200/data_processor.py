"""DataProcessor.transform_points_to_voxels with the reference's config wiring
(pcdet/datasets/processor/data_processor.py:8-17, 43-81), computed by the sm_100a voxelizer.

Only the voxelization hook of the reference's DataProcessor is on the hot path; the other
hooks (range mask, shuffle, cylinder transform, point sampling) stay with the caller's dataset code.  The hook
keeps the reference's two-phase protocol: called with ``config`` only it builds the generator from

    config.VOXEL_SIZE, config.MAX_POINTS_PER_VOXEL, config.MAX_NUMBER_OF_VOXELS[mode]   (mode = 'train' | 'test')

sets ``grid_size`` / ``voxel_size`` on the processor and returns the bound per-sample function; called with a
``data_dict`` it writes ``voxels``, ``voxel_coords`` (z,y,x) and ``voxel_num_points``, dropping the xyz columns of
``voxels`` when ``data_dict['use_lead_xyz']`` is false (data_processor.py:72-73).  numpy in, numpy out like the
reference (the arrays come back from the device); CUDA tensors in, CUDA tensors out.
"""
from functools import partial

import numpy as np

from .voxel_generator import VoxelGenerator


def _get(cfg, name):
    """Config access for EasyDict / dict / attribute objects alike (the reference uses EasyDict, config.py:84)."""
    if isinstance(cfg, dict):
        return cfg[name]
    return getattr(cfg, name)


class DataProcessor(object):
    """processor_configs: list of configs with a NAME key; only 'transform_points_to_voxels' is built here, any
    other NAME raises NotImplementedError (out of the hot path's scope) unless ``passthrough`` maps it to a callable
    ``fn(data_dict, config) -> data_dict`` supplied by the caller."""

    def __init__(self, processor_configs, point_cloud_range, training, passthrough=None, device=None):
        self.point_cloud_range = np.asarray(point_cloud_range)
        self.training = training
        self.mode = 'train' if training else 'test'
        self.grid_size = self.voxel_size = None
        self.device = device
        self.data_processor_queue = []
        passthrough = passthrough or {}
        for cur_cfg in processor_configs:
            name = _get(cur_cfg, 'NAME')
            if name == 'transform_points_to_voxels':
                cur_processor = self.transform_points_to_voxels(config=cur_cfg)
            elif name in passthrough:
                cur_processor = partial(passthrough[name], config=cur_cfg)
            else:
                raise NotImplementedError("DataProcessor hook %r is outside the voxelize+backbone hot path; pass it "
                                          "through `passthrough={%r: fn}`" % (name, name))
            self.data_processor_queue.append(cur_processor)

    def transform_points_to_voxels(self, data_dict=None, config=None, voxel_generator=None):
        if data_dict is None:
            voxel_generator = VoxelGenerator(
                voxel_size=_get(config, 'VOXEL_SIZE'),
                point_cloud_range=self.point_cloud_range,
                max_num_points=_get(config, 'MAX_POINTS_PER_VOXEL'),
                max_voxels=_get(_get(config, 'MAX_NUMBER_OF_VOXELS'), self.mode),
                device=self.device,
            )
            grid_size = (self.point_cloud_range[3:6] - self.point_cloud_range[0:3]) / np.array(_get(config, 'VOXEL_SIZE'))
            self.grid_size = np.round(grid_size).astype(np.int64)
            self.voxel_size = _get(config, 'VOXEL_SIZE')
            return partial(self.transform_points_to_voxels, voxel_generator=voxel_generator)

        points = data_dict['points']
        voxels, coordinates, num_points = voxel_generator.generate(points)
        if not data_dict['use_lead_xyz']:
            voxels = voxels[..., 3:]  # remove xyz in voxels(N, 3)
        data_dict['voxels'] = voxels
        data_dict['voxel_coords'] = coordinates
        data_dict['voxel_num_points'] = num_points
        return data_dict

    def forward(self, data_dict):
        """data_processor.py:144-155: runs the queue in order."""
        for cur_processor in self.data_processor_queue:
            data_dict = cur_processor(data_dict=data_dict)
        return data_dict
