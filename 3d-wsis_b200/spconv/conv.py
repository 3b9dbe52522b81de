"""Sparse convolution modules: contract of modules/lib/spconv/spconv/conv.py:51-355.

Same constructor signatures, same `weight` parameter of shape [*kernel_size, in_channels, out_channels]
(conv.py:98-99) initialised with kaiming_uniform_(a=sqrt(5)) in the same RNG order (conv.py:106-112), same
`indice_dict[indice_key] = (outids, indices, indice_pairs, indice_pair_num, spatial_shape)` cache contract
(conv.py:152) -- so released checkpoints load and fixed-seed models match.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

import spconv
import spconv.functional as Fsp
from spconv import ops
from spconv.modules import SparseModule
from wsis_b200 import ops as W


def _calculate_fan_in_and_fan_out_hwio(tensor):
    dimensions = tensor.ndimension()
    if dimensions < 2:
        raise ValueError("Fan in and fan out can not be computed for tensor with fewer than 2 dimensions")
    if dimensions == 2:
        return tensor.size(-2), tensor.size(-1)
    receptive = tensor[..., 0, 0].numel()
    return tensor.size(-2) * receptive, tensor.size(-1) * receptive


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None):
        super(SparseConvolution, self).__init__()
        assert groups == 1
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * ndim  # noqa: E731
        kernel_size, stride, padding = as_list(kernel_size), as_list(stride), as_list(padding)
        dilation, output_padding = as_list(dilation), as_list(output_padding)
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
        self.weight = Parameter(torch.Tensor(*kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        self._packed = W.PackedWeights()  # pre-swizzled bf16 image of `weight` for the tcgen05 kernel

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = _calculate_fan_in_and_fan_out_hwio(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input, _prologue=None, _residual=None, _bn_train=None):
        """`_prologue` = (scale, shift, relu) and `_residual` are the fusion hooks used by SparseSequential and by
        wsis_b200.model (inference only); plain calls behave exactly like the reference module."""
        assert isinstance(input, spconv.SparseConvTensor)
        features = input.features
        device = features.device
        indices = input.indices
        spatial_shape = input.spatial_shape
        batch_size = input.batch_size
        if not self.subm:
            if self.transposed:
                out_spatial_shape = ops.get_deconv_output_size(spatial_shape, self.kernel_size, self.stride,
                                                               self.padding, self.dilation, self.output_padding)
            else:
                out_spatial_shape = ops.get_conv_output_size(spatial_shape, self.kernel_size, self.stride,
                                                             self.padding, self.dilation)
        else:
            out_spatial_shape = spatial_shape
        if self.conv1x1:
            assert _prologue is None and _residual is None
            out1 = None
            if not torch.is_grad_enabled() and not self.training:
                # inference: the identity rulebook on the tensor-core kernel instead of a library GEMM
                out1 = W.dense_rows(features, self.weight.view(self.in_channels, self.out_channels), packed=self._packed,
                                    holder=indices)
            input.features = out1 if out1 is not None else \
                torch.mm(input.features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                input.features += self.bias
            return input
        datas = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert datas is not None and self.indice_key is not None
            _, outids, indice_pairs, indice_pair_num, out_spatial_shape = datas
            assert indice_pairs.shape[0] == np.prod(self.kernel_size), \
                "inverse conv must have same kernel size as its couple conv"
        else:
            if self.indice_key is not None and datas is not None:
                outids, _, indice_pairs, indice_pair_num, _ = datas
            else:
                outids, indice_pairs, indice_pair_num = ops.get_indice_pairs(
                    indices, batch_size, spatial_shape, self.kernel_size, self.stride, self.padding, self.dilation,
                    self.output_padding, self.subm, self.transposed, grid=input.grid)
                input.indice_dict[self.indice_key] = (outids, indices, indice_pairs, indice_pair_num, spatial_shape)
        needs_grad = torch.is_grad_enabled() and (features.requires_grad or self.weight.requires_grad)
        if _bn_train is not None:
            # training: batch-statistics BatchNorm + ReLU ride in this conv's gather prologue (forward and wgrad), the
            # backward is dgrad + one reduction + one elementwise pass (wsis_b200/train.py)
            from wsis_b200 import train as T
            n_feat = features.shape[0]
            if self.inverse:
                rb = ops._rulebook_of(indice_pairs, indice_pair_num, outids.shape[0], n_feat, False)
            else:
                rb = ops._rulebook_of(indice_pairs, indice_pair_num, n_feat, outids.shape[0], self.subm)
            out_features = T.bn_relu_conv(features, _bn_train, self.weight, rb, "inv" if self.inverse else "fwd",
                                          residual=_residual, packed=self._packed)
        elif needs_grad:
            assert _prologue is None and _residual is None, "fusion hooks are inference-only"
            fn = Fsp.indice_subm_conv if self.subm else (Fsp.indice_inverse_conv if self.inverse else Fsp.indice_conv)
            out_features = fn(features, self.weight, indice_pairs.to(device), indice_pair_num, outids.shape[0])
        else:
            out_features = ops.indice_conv(features, self.weight, indice_pairs, indice_pair_num, outids.shape[0],
                                           self.inverse, self.subm, _prologue=_prologue, _residual=_residual,
                                           _packed=self._packed)
        if self.bias is not None:
            out_features += self.bias
        out_tensor = spconv.SparseConvTensor(out_features, outids, out_spatial_shape, batch_size)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        return out_tensor


def _make(ndim, name, doc, **fixed):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        SparseConvolution.__init__(self, ndim, in_channels, out_channels, kernel_size, stride, padding, dilation,
                                   groups, bias, indice_key=indice_key, **fixed)
    return type(name, (SparseConvolution,), {"__init__": __init__, "__doc__": doc})


SparseConv2d = _make(2, "SparseConv2d", "conv.py:177-198")
SparseConv3d = _make(3, "SparseConv3d", "conv.py:201-222")
SparseConvTranspose2d = _make(2, "SparseConvTranspose2d", "conv.py:224-245", transposed=True)
SparseConvTranspose3d = _make(3, "SparseConvTranspose3d", "conv.py:248-271", transposed=True)
SubMConv2d = _make(2, "SubMConv2d", "conv.py:308-330", subm=True)
SubMConv3d = _make(3, "SubMConv3d", "conv.py:333-355", subm=True)


def _make_inverse(ndim, name, doc):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        SparseConvolution.__init__(self, ndim, in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                                   indice_key=indice_key)
    return type(name, (SparseConvolution,), {"__init__": __init__, "__doc__": doc})


SparseInverseConv2d = _make_inverse(2, "SparseInverseConv2d", "conv.py:274-288")
SparseInverseConv3d = _make_inverse(3, "SparseInverseConv3d", "conv.py:291-305")
