"""SparseModule / SparseSequential: contract of modules/lib/spconv/spconv/modules.py:40-130.

SparseSequential keeps the reference's routing rule (sparse modules get the SparseConvTensor, dense nn modules
are applied to `.features` in place on the caller's object, :118-130).  One addition that does not change
results: in inference (no grad, module in eval mode) the pattern  BatchNorm1d -> ReLU -> sparse conv  that every
ResidualBlock / UBlock stage uses (sparse_unet3d.py:163-172,258-298) is executed as ONE kernel, the folded
BatchNorm affine and the ReLU becoming the conv's gather prologue.
"""
import sys
from collections import OrderedDict

import torch
from torch import nn

import spconv
from wsis_b200 import ops as wsis_ops


class SparseModule(nn.Module):
    """Marker base class: subclasses receive a SparseConvTensor inside SparseSequential.

    Derived parameter images (packed conv weights, folded BatchNorm) are invalidated whenever parameters are
    (re)loaded, moved or converted, or the train/eval mode changes (wsis_b200.ops.invalidate_caches)."""

    def _load_from_state_dict(self, *args, **kwargs):
        wsis_ops.invalidate_caches()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        wsis_ops.invalidate_caches()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        wsis_ops.invalidate_caches()
        return super().train(mode)


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def _fold_bn(bn):
    """Eval-mode BatchNorm1d as (scale, shift); cached on the module per parameter version."""
    key = (bn.weight._version if bn.weight is not None else -1, bn.bias._version if bn.bias is not None else -1,
           bn.running_mean._version, bn.running_var._version, bn.running_mean.data_ptr(), wsis_ops.cache_epoch())
    cache = getattr(bn, "_wsis_fold", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            inv = torch.rsqrt(bn.running_var.float() + bn.eps)
            w = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
            b = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
            scale = (w * inv).contiguous()
            shift = (b - bn.running_mean.float() * scale).contiguous()
        cache = (key, scale, shift)
        bn._wsis_fold = cache
    return cache[1], cache[2]


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super(SparseSequential, self).__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if sys.version_info < (3, 6):
                raise ValueError("kwargs only supported in py36+")
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self._sparity_dict = {}

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError('index {} is out of range'.format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    @property
    def sparity_dict(self):
        return self._sparity_dict

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def _fusable(self, mods, i, input):
        """BatchNorm1d(eval) -> ReLU -> sparse conv (not 1x1), no autograd, CUDA fp32 features."""
        if i + 2 >= len(mods) or torch.is_grad_enabled():
            return False
        bn, act, conv = mods[i], mods[i + 1], mods[i + 2]
        return (isinstance(bn, nn.BatchNorm1d) and not bn.training and bn.track_running_stats
                and type(act) is nn.ReLU and isinstance(conv, spconv.conv.SparseConvolution) and not conv.conv1x1
                and isinstance(input, spconv.SparseConvTensor) and input.features.is_cuda
                and input.features.dtype == torch.float32 and input.indices.shape[0] != 0)

    def _fusable_tail(self, mods, i, input):
        if i + 1 >= len(mods) or torch.is_grad_enabled():
            return False
        bn, act = mods[i], mods[i + 1]
        return (isinstance(bn, nn.BatchNorm1d) and not bn.training and bn.track_running_stats
                and type(act) is nn.ReLU and isinstance(input, spconv.SparseConvTensor) and input.features.is_cuda
                and input.features.dtype == torch.float32 and input.indices.shape[0] != 0)

    def _trainable(self, mods, i, input):
        """BatchNorm1d(training) -> ReLU [-> sparse conv (not 1x1)] under autograd on CUDA fp32 features: 2 = with
        the conv behind it, 1 = BatchNorm + ReLU only, 0 = no."""
        if i + 1 >= len(mods) or not torch.is_grad_enabled():
            return 0
        from wsis_b200 import train as T
        if not T.FUSED:
            return 0
        bn, act = mods[i], mods[i + 1]
        if not (isinstance(bn, nn.BatchNorm1d) and bn.training and type(act) is nn.ReLU
                and isinstance(input, spconv.SparseConvTensor) and input.features.is_cuda
                and input.features.dtype == torch.float32 and input.indices.shape[0] > 1):
            return 0
        if i + 2 < len(mods) and isinstance(mods[i + 2], spconv.conv.SparseConvolution) and not mods[i + 2].conv1x1 \
                and mods[i + 2].bias is None:
            return 2
        return 1

    def forward(self, input):
        mods = list(self._modules.items())
        vals = [m for _, m in mods]
        i = 0
        while i < len(mods):
            k, module = mods[i]
            tr = self._trainable(vals, i, input)
            if tr == 2:
                self._sparity_dict[mods[i + 2][0]] = input.sparity
                input = vals[i + 2](input, _bn_train=module)
                i += 3
                continue
            if tr == 1:
                from wsis_b200 import train as T
                input.features = T.batch_norm_train(input.features, module, True)
                i += 2
                continue
            if self._fusable(vals, i, input):
                scale, shift = _fold_bn(module)
                self._sparity_dict[mods[i + 2][0]] = input.sparity
                input = vals[i + 2](input, _prologue=(scale, shift, 1))
                i += 3
                continue
            if self._fusable_tail(vals, i, input):
                # BatchNorm1d(eval) -> ReLU with no conv behind it (output_layer, backbone_3D_WSIS.py:52-55)
                scale, shift = _fold_bn(module)
                input.features = wsis_ops.affine_relu(input.features, scale, shift, True)
                i += 2
                continue
            if is_spconv_module(module):
                assert isinstance(input, spconv.SparseConvTensor)
                self._sparity_dict[k] = input.sparity
                input = module(input)
            else:
                if isinstance(input, spconv.SparseConvTensor):
                    if input.indices.shape[0] != 0:
                        input.features = module(input.features)
                else:
                    input = module(input)
            i += 1
        return input
