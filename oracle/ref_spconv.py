"""TEST INFRASTRUCTURE ONLY -- loader for oracle/_ref/libspconv_ref.so, i.e. the UNMODIFIED reference
spconv 1.0 CPU path compiled from /root/reference by oracle/Makefile (the .so travels to the GPU box,
the reference sources do not).  It registers torch.ops.spconv.{get_indice_pairs_3d, indice_conv_fp32,
indice_conv_backward_fp32} exactly as the reference's Python layer calls them
(modules/lib/spconv/spconv/ops.py:85-89,109-112,124-126).

Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libspconv_ref.so")
_loaded = False


def available():
    return os.path.exists(_SO)


def load():
    global _loaded
    if not _loaded:
        if not available():
            raise FileNotFoundError(_SO + " (run `make -C oracle ref` where /root/reference exists)")
        torch.ops.load_library(_SO)
        _loaded = True
    return torch.ops.spconv


def _l3(v):
    return list(v) if isinstance(v, (list, tuple)) else [int(v)] * 3


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     out_padding=0, subm=False, transpose=False):
    """Mirrors ops.get_indice_pairs, spconv/ops.py:45-98 (CPU tensors in, CPU tensors out)."""
    ops = load()
    ksize, stride, padding, dilation, out_padding = map(_l3, (ksize, stride, padding, dilation, out_padding))
    spatial_shape = [int(s) for s in spatial_shape]
    if subm:
        out_shape = spatial_shape
    else:
        out_shape = [(spatial_shape[i] + 2 * padding[i] - dilation[i] * (ksize[i] - 1) - 1) // stride[i] + 1
                     for i in range(3)]
    return ops.get_indice_pairs_3d(indices.int().contiguous(), int(batch_size), out_shape, spatial_shape, ksize,
                                   stride, padding, dilation, out_padding, int(subm), int(transpose))


def indice_conv(features, filters, pairs, num, n_out, inverse=False, subm=False):
    return load().indice_conv_fp32(features, filters, pairs, num, int(n_out), int(inverse), int(subm))


def indice_conv_backward(features, filters, dout, pairs, num, inverse=False, subm=False):
    return load().indice_conv_backward_fp32(features, filters, dout, pairs, num, int(inverse), int(subm))
