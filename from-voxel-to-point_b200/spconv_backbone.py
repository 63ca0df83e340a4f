"""VoxelBackBone8x / VoxelResBackBone8x with the reference's constructors, module names (= checkpoint
keys), config keys and batch_dict contract (pcdet/models/backbones_3d/spconv_backbone.py:10-290).

Two execution paths, same numbers:
  * eval mode on CUDA -> engine.BackboneEngine: all rulebooks first, then one fused
    conv+bias+BatchNorm(+residual)+ReLU kernel per layer, one host read of the row counts at the end;
  * training mode (or FUSED: False in model_cfg) -> the plain module graph below, which mirrors the
    reference op for op (separate BatchNorm1d / ReLU modules on ``.features``).

Extra, optional model_cfg keys (absent in the reference's YAML, defaults keep its behaviour):
  PRECISION: 'fp32' (default) | 'bf16';  FUSED: True;  MATERIALIZE_PAIRS: True | 'lazy' | False (fill indice_dict with the
  reference-layout pair tensors; the fused kernels themselves only need the neighbour map);  SORT_ROWS: True
  (process output rows in neighbour-mask order inside the tensor-core kernels; results are unchanged).
"""
from functools import partial

import torch
import torch.nn as nn

from . import spconv
from .batchnorm import BatchNorm1d
from .engine import BackboneEngine


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0, conv_type='subm',
                   norm_fn=None):
    if conv_type == 'subm':
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == 'spconv':
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                   bias=False, indice_key=indice_key)
    elif conv_type == 'inverseconv':
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
    else:
        raise NotImplementedError
    return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


class SparseBasicBlock(spconv.SparseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_fn=None, downsample=None, indice_key=None):
        super(SparseBasicBlock, self).__init__()
        assert norm_fn is not None
        bias = norm_fn is not None
        self.conv1 = spconv.SubMConv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                                       indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                                       indice_key=indice_key)
        self.bn2 = norm_fn(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.conv1(x)
        out.features = self.bn1(out.features)
        out.features = self.relu(out.features)
        out = self.conv2(out)
        out.features = self.bn2(out.features)
        if self.downsample is not None:
            identity = self.downsample(x)
        out.features += identity.features
        out.features = self.relu(out.features)
        return out


def _detach_outputs(outs):
    """Copies of the engine's exported tensors (features, indices, rulebook tensors), sharing one indice_dict /
    nbr_dict like the originals do."""
    memo = {}

    def cp(t):
        if t is None or not isinstance(t, torch.Tensor):
            return t
        key = (t.data_ptr(), tuple(t.shape), t.dtype)
        if key not in memo:
            memo[key] = t.clone()
        return memo[key]

    indice_dict = nbr_dict = None
    res = {}
    for name, sp in outs.items():
        if indice_dict is None:
            indice_dict = {k: (cp(v[0]), cp(v[1]), cp(v[2]), cp(v[3]), v[4]) for k, v in sp.indice_dict.items()}
            nbr_dict = {k: cp(v) for k, v in sp.nbr_dict.items()}
        t = spconv.SparseConvTensor(cp(sp.features), cp(sp.indices), sp.spatial_shape, sp.batch_size)
        t.indice_dict, t.nbr_dict = indice_dict, nbr_dict
        res[name] = t
    return res


class _BackboneBase(nn.Module):
    """Shared forward: engine in eval mode, module graph otherwise."""

    def _cfg(self, key, default):
        try:
            return self.model_cfg.get(key, default)
        except AttributeError:
            return default

    def _modules_forward(self, input_sp_tensor):
        x = self.conv_input(input_sp_tensor)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        return dict(x_conv1=x_conv1, x_conv2=x_conv2, x_conv3=x_conv3, x_conv4=x_conv4, out=out)

    def make_engine(self, materialize_pairs=None):
        """A new BackboneEngine over this module tree with the options of model_cfg (PRECISION, MATERIALIZE_PAIRS,
        SORT_ROWS, CAP_GROWTH).  materialize_pairs overrides the config when the config does not set it."""
        mp = self._cfg('MATERIALIZE_PAIRS', True if materialize_pairs is None else materialize_pairs)
        return BackboneEngine(self, precision=self._cfg('PRECISION', 'fp32'),
                              materialize_pairs=mp if mp in (True, False, 'lazy') else bool(mp),
                              sort_rows=self._cfg('SORT_ROWS', True),
                              cap_growth=self._cfg('CAP_GROWTH', 'auto'))

    def get_engine(self):
        precision = self._cfg('PRECISION', 'fp32')
        eng = getattr(self, '_engine', None)
        if eng is None or eng.precision != precision:
            eng = self.make_engine()
            object.__setattr__(self, '_engine', eng)  # not a submodule, never in the state_dict
        return eng

    def forward(self, batch_dict):
        """batch_dict: batch_size, voxel_features [N,C], voxel_coords [N,4] (batch,z,y,x) -> adds
        encoded_spconv_tensor(+_stride), multi_scale_3d_features, multi_scale_3d_strides."""
        voxel_features, voxel_coords = batch_dict['voxel_features'], batch_dict['voxel_coords']
        batch_size = batch_dict['batch_size']
        coords = voxel_coords.int()
        # The fused engine is an inference transform (BatchNorm folded, no autograd graph): it is used in eval mode
        # when nothing can ask for gradients, i.e. under torch.no_grad() like the reference's eval loop
        # (tools/eval_utils/eval_utils.py:58) or when neither the input nor any parameter requires grad.  Eval mode
        # with gradients enabled (frozen-BN fine-tuning) takes the differentiable module path.
        wants_grad = torch.is_grad_enabled() and (voxel_features.requires_grad or
                                                  any(p.requires_grad for p in self.parameters()))
        fused = (not self.training) and bool(self._cfg('FUSED', True)) and voxel_features.is_cuda and not wants_grad
        if fused and getattr(self, '_engine_unavailable', False):
            fused = False
        if fused:
            try:
                self.get_engine()
            except NotImplementedError as e:
                # a module tree the engine's tracer does not recognise (edited backbone): say so once and run the module
                # graph, whose convolutions still use the tensor-core kernels when no gradients are needed
                import warnings
                warnings.warn("fv2p_b200: %s; running the module graph instead of the fused engine" % e)
                object.__setattr__(self, '_engine_unavailable', True)
                fused = False
        if fused:
            with torch.no_grad():
                outs = self.get_engine()(voxel_features.float().contiguous(), coords.contiguous(), batch_size)
            if not bool(self._cfg('ALIAS_OUTPUTS', False)):
                # The engine hands out views into its grow-only arena, which the next forward() overwrites; the
                # reference returns fresh tensors, so the module API copies them out (11 MB for a KITTI batch of 8).
                # ALIAS_OUTPUTS: True keeps the zero-copy views (HotPath does, it documents their lifetime).
                outs = _detach_outputs(outs)
        else:
            input_sp_tensor = spconv.SparseConvTensor(features=voxel_features, indices=coords,
                                                      spatial_shape=self.sparse_shape, batch_size=batch_size)
            outs = self._modules_forward(input_sp_tensor)
        batch_dict.update({'encoded_spconv_tensor': outs['out'], 'encoded_spconv_tensor_stride': 8})
        batch_dict.update({
            'multi_scale_3d_features': {k: outs[k] for k in ('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4')},
            'multi_scale_3d_strides': {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8},
        })
        return batch_dict


class VoxelBackBone8x(_BackboneBase):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        norm_fn = partial(BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = grid_size[::-1] + [1, 0, 0]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key='subm1'),
            norm_fn(16),
            nn.ReLU(),
        )
        block = post_act_block
        self.conv1 = spconv.SparseSequential(
            block(16, 16, 3, norm_fn=norm_fn, padding=1, indice_key='subm1'),
        )
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv2', conv_type='spconv'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key='subm2'),
        )
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv3', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm3'),
        )
        self.conv4 = spconv.SparseSequential(
            block(64, 64, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key='spconv4', conv_type='spconv'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key='subm4'),
        )
        last_pad = self._cfg('last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key='spconv_down2'),
            norm_fn(128),
            nn.ReLU(),
        )
        # a dict in this fork (spconv_backbone.py:124-131)
        self.num_point_features = {'x_conv1': 16, 'x_conv2': 32, 'x_conv3': 64, 'x_conv4': 64}


class VoxelResBackBone8x(_BackboneBase):
    def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        norm_fn = partial(BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = grid_size[::-1] + [1, 0, 0]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key='subm1'),
            norm_fn(16),
            nn.ReLU(),
        )
        block = post_act_block
        self.conv1 = spconv.SparseSequential(
            SparseBasicBlock(16, 16, norm_fn=norm_fn, indice_key='res1'),
            SparseBasicBlock(16, 16, norm_fn=norm_fn, indice_key='res1'),
        )
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv2', conv_type='spconv'),
            SparseBasicBlock(32, 32, norm_fn=norm_fn, indice_key='res2'),
            SparseBasicBlock(32, 32, norm_fn=norm_fn, indice_key='res2'),
        )
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key='spconv3', conv_type='spconv'),
            SparseBasicBlock(64, 64, norm_fn=norm_fn, indice_key='res3'),
            SparseBasicBlock(64, 64, norm_fn=norm_fn, indice_key='res3'),
        )
        self.conv4 = spconv.SparseSequential(
            block(64, 128, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key='spconv4',
                  conv_type='spconv'),
            SparseBasicBlock(128, 128, norm_fn=norm_fn, indice_key='res4'),
            SparseBasicBlock(128, 128, norm_fn=norm_fn, indice_key='res4'),
        )
        last_pad = self._cfg('last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(128, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key='spconv_down2'),
            norm_fn(128),
            nn.ReLU(),
        )
        self.num_point_features = 128
