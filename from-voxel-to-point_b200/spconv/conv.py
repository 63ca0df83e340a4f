"""SparseConvolution / SubMConv3d / SparseConv3d with the reference's constructor signatures, parameter
names and shapes (pcdet/ops/spconv/conv.py:48-229, 258-281, 429-453), so reference checkpoints load
unchanged: weight [kD,kH,kW,Cin,Cout], optional bias [Cout]."""
import math

import numpy as np
import torch
from torch.nn import init
from torch.nn.parameter import Parameter

from .. import _lib
from . import functional as Fsp
from . import ops
from .modules import SparseModule
from .structure import SparseConvTensor


def _calculate_fan_in_and_fan_out_hwio(tensor):
    dimensions = tensor.ndimension()
    if dimensions < 2:
        raise ValueError('fan in and fan out can not be computed for tensor with fewer than 2 dimensions')
    if dimensions == 2:
        fan_in, fan_out = tensor.size(-2), tensor.size(-1)
    else:
        receptive_field_size = tensor[..., 0, 0].numel()
        fan_in = tensor.size(-2) * receptive_field_size
        fan_out = tensor.size(-1) * receptive_field_size
    return fan_in, fan_out


class SparseConvolution(SparseModule):

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False):
        super(SparseConvolution, self).__init__()
        assert groups == 1
        if not isinstance(kernel_size, (list, tuple)):
            kernel_size = [kernel_size] * ndim
        if not isinstance(stride, (list, tuple)):
            stride = [stride] * ndim
        if not isinstance(padding, (list, tuple)):
            padding = [padding] * ndim
        if not isinstance(dilation, (list, tuple)):
            dilation = [dilation] * ndim
        if not isinstance(output_padding, (list, tuple)):
            output_padding = [output_padding] * ndim
        for d, s in zip(dilation, stride):
            assert any([s == 1, d == 1]), "don't support this."
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.conv1x1 = np.prod(kernel_size) == 1
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = output_padding
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.fused_bn = fused_bn
        self.weight = Parameter(torch.Tensor(*kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = _calculate_fan_in_and_fan_out_hwio(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        features = input.features
        indices = input.indices
        spatial_shape = input.spatial_shape
        batch_size = input.batch_size
        if self.transposed or self.inverse:
            raise NotImplementedError("transposed / inverse sparse convolution is outside the hot path")
        if not self.subm:
            out_spatial_shape = ops.get_conv_output_size(spatial_shape, self.kernel_size, self.stride, self.padding,
                                                         self.dilation)
        else:
            out_spatial_shape = spatial_shape
        if self.conv1x1:  # conv.py:137-148
            features = torch.mm(input.features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                features += self.bias
            out_tensor = SparseConvTensor(features, input.indices, input.spatial_shape, input.batch_size)
            out_tensor.indice_dict = input.indice_dict
            out_tensor.nbr_dict = input.nbr_dict
            out_tensor.grid = input.grid
            return out_tensor
        datas = input.find_indice_pair(self.indice_key)
        nbr = None
        if self.indice_key is not None and datas is not None:
            outids, _, indice_pairs, indice_pair_num, _ = datas
            nbr = input.nbr_dict.get(self.indice_key)
        else:
            outids, indice_pairs, indice_pair_num, nbr = ops.get_indice_pairs(
                indices, batch_size, spatial_shape, self.kernel_size, self.stride, self.padding, self.dilation,
                self.output_padding, self.subm, self.transposed, grid=input.grid, return_nbr=True)
            input.indice_dict[self.indice_key] = (outids, indices, indice_pairs, indice_pair_num, spatial_shape)
            if self.indice_key is not None:
                input.nbr_dict[self.indice_key] = nbr
        out_features = self._tensor_core_forward(input, features, outids, indice_pairs, indice_pair_num, nbr)
        if out_features is None:
            if self.subm:
                out_features = Fsp.indice_subm_conv(features, self.weight, indice_pairs, indice_pair_num,
                                                    outids.shape[0], nbr)
            else:
                out_features = Fsp.indice_conv(features, self.weight, indice_pairs, indice_pair_num, outids.shape[0],
                                               nbr)
            if self.bias is not None:
                out_features += self.bias
        out_tensor = SparseConvTensor(out_features, outids, out_spatial_shape, batch_size)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.nbr_dict = input.nbr_dict
        out_tensor.grid = input.grid
        return out_tensor


    def _tensor_core_forward(self, input, features, outids, indice_pairs, indice_pair_num, nbr):
        """Inference through the module API (any module graph, FUSED: False, trees the engine does not recognise): when
        nothing can ask for gradients and the shape fits, the tcgen05 kernel runs here too - fp32 features through the
        fp32 tensor-core mode (split-bf16 products), bf16 features through the bf16 mode, bias fused - on the rulebook's grouped row
        order, which is built once per indice_key and cached next to the neighbour map.  Returns None when the plain
        path (fp32 FMA + autograd) has to be taken."""
        if torch.is_grad_enabled() and (features.requires_grad or self.weight.requires_grad or
                                        (self.bias is not None and self.bias.requires_grad)):
            return None
        if not features.is_cuda or features.dtype not in (torch.float32, torch.bfloat16):
            return None
        mode = _lib.MODE_TF32X3_TC if features.dtype == torch.float32 else _lib.MODE_BF16_TC
        kvol = int(np.prod(self.kernel_size))
        if _lib.load().fv2p_pack_weight_bytes(kvol, self.in_channels, self.out_channels, mode) == 0:
            return None
        n_out = int(outids.shape[0])
        if n_out == 0:
            return None
        if nbr is None:
            nbr = ops.pairs_to_nbr(indice_pairs, indice_pair_num, n_out, False)
        gkey = ("grouped", self.indice_key)
        grouped = input.nbr_dict.get(gkey) if self.indice_key is not None else None
        if grouped is None:
            perm, nbr_sorted, order = ops.sort_rows_by_mask(nbr.contiguous(), n_out, return_tile_order=True)
            grouped = (perm.contiguous(), nbr_sorted.contiguous(), order.contiguous())
            if self.indice_key is not None:
                input.nbr_dict[gkey] = grouped
        wkey = (self.weight.data_ptr(), self.weight._version, mode)
        cached = getattr(self, "_packed", None)
        if cached is None or cached[0] != wkey:
            w = self.weight.detach().reshape(kvol, self.in_channels, self.out_channels).float().contiguous()
            cached = (wkey, ops.pack_weight(w, mode))
            object.__setattr__(self, "_packed", cached)
        bias = self.bias.detach().float().contiguous() if self.bias is not None else None
        return ops.conv_forward(features.contiguous(), cached[1], grouped[1], n_out, bias=bias, mode=mode,
                                row_perm=grouped[0], tile_order=grouped[2])


class SparseConv3d(SparseConvolution):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super(SparseConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation,
                                           groups, bias, indice_key=indice_key)


class SubMConv3d(SparseConvolution):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super(SubMConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation,
                                         groups, bias, True, indice_key=indice_key)
