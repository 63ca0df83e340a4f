"""Autograd wrappers with the reference's names (pcdet/ops/spconv/functional.py:20-101)."""
from torch.autograd import Function

from . import ops


class SparseConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out, nbr=None):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, False, False, nbr)

    @staticmethod
    def backward(ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output.contiguous(), indice_pairs,
                                                        indice_pair_num, False)
        return input_bp, filters_bp, None, None, None, None


class SubMConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out, nbr=None):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, False, True, nbr)

    @staticmethod
    def backward(ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output.contiguous(), indice_pairs,
                                                        indice_pair_num, False, True)
        return input_bp, filters_bp, None, None, None, None


indice_conv = SparseConvFunction.apply
indice_subm_conv = SubMConvFunction.apply
