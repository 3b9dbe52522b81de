"""Autograd entry points: contract of modules/lib/spconv/spconv/functional.py:21-117."""
from torch.autograd import Function

import spconv.ops as ops


class _ConvFunction(Function):
    inverse = False
    subm = False

    @classmethod
    def _fwd(cls, ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(indice_pairs, indice_pair_num, features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, cls.inverse, cls.subm)

    @classmethod
    def _bwd(cls, ctx, grad_output):
        indice_pairs, indice_pair_num, features, filters = ctx.saved_tensors
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output.contiguous(), indice_pairs,
                                                        indice_pair_num, cls.inverse, cls.subm,
                                                        _need_input_grad=ctx.needs_input_grad[0])
        return input_bp, filters_bp, None, None, None


class SparseConvFunction(_ConvFunction):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SparseConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out)

    @staticmethod
    def backward(ctx, grad_output):
        return SparseConvFunction._bwd(ctx, grad_output)


class SparseInverseConvFunction(_ConvFunction):
    inverse = True

    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SparseInverseConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out)

    @staticmethod
    def backward(ctx, grad_output):
        return SparseInverseConvFunction._bwd(ctx, grad_output)


class SubMConvFunction(_ConvFunction):
    subm = True

    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        return SubMConvFunction._fwd(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out)

    @staticmethod
    def backward(ctx, grad_output):
        return SubMConvFunction._bwd(ctx, grad_output)


indice_conv = SparseConvFunction.apply
indice_inverse_conv = SparseInverseConvFunction.apply
indice_subm_conv = SubMConvFunction.apply
