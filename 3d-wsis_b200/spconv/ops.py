"""Functional layer: contract of modules/lib/spconv/spconv/ops.py:19-135 (same names, argument meaning and
error behaviour), implemented over wsis_b200.ops instead of torch.ops.spconv.*.

The returned `indice_pairs` tensor is the reference-format int32 [K,2,N] (-1 padded) rulebook; it additionally
carries the output-stationary neighbour maps the kernels actually consume (attribute `_wsis_rulebook`).  Pairs
supplied by a caller without that attribute are converted on first use.
"""
import torch

from wsis_b200 import ops as W


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        size = (input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
        output_size.append(1 if kernel_size[i] == -1 else size)
    return output_size


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation, output_padding):
    ndim = len(input_size)
    output_size = []
    for i in range(ndim):
        if kernel_size[i] == -1:
            raise ValueError("deconv don't support kernel_size < 0")
        output_size.append((input_size[i] - 1) * stride[i] - 2 * padding[i] + kernel_size[i] + output_padding[i])
    return output_size


def _listify(v, ndim):
    return list(v) if isinstance(v, (list, tuple)) else [v] * ndim


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1, out_padding=0,
                     subm=False, transpose=False, grid=None):
    """-> (outids int32[M,4], indice_pairs int32[K,2,N], indice_pair_num int32[K]);  `grid` is accepted and
    ignored (the hash-table builder needs no dense grid)."""
    ndim = indices.shape[1] - 1
    if ndim != 3:
        raise NotImplementedError("wsis_b200 builds 3-D rulebooks only (the 3D-WSIS hot path)")
    ksize, stride, padding, dilation, out_padding = (_listify(v, ndim) for v in
                                                     (ksize, stride, padding, dilation, out_padding))
    for d, s in zip(dilation, stride):
        assert any([s == 1, d == 1]), "don't support this."
    if transpose:
        raise NotImplementedError("transposed sparse conv is not on the 3D-WSIS hot path (SURVEY.md 2b)")
    if indices.dtype != torch.int32:
        indices = indices.int()
    if subm:
        rb = W.rulebook_subm(indices, spatial_shape, ksize, dilation, batch_size)
    else:
        rb, _ = W.rulebook_conv(indices, spatial_shape, ksize, stride, padding, dilation, batch_size)
    if _LAZY_PAIRS:
        lazy = LazyPairs(rb)
        return rb.out_coords, lazy, lazy
    pairs, num = rb.pairs()
    pairs._wsis_rulebook = rb
    return rb.out_coords, pairs, num


# The reference-format (indice_pairs, indice_pair_num) tensors exist only because indice_dict stores them
# (conv.py:152); no kernel of this package reads them.  Inside `with lazy_pairs():` (wsis_b200.pipeline: the fused
# inference path, where nothing looks into indice_dict) get_indice_pairs returns a LazyPairs handle instead and the
# [K, 2, N] tensor (117 MB at the first U-Net level of a 4-scene batch) is only written if somebody asks for it.
_LAZY_PAIRS = False


class LazyPairs(object):
    def __init__(self, rb):
        self._wsis_rulebook = rb

    def tensors(self):
        return self._wsis_rulebook.pairs()

    @property
    def shape(self):
        rb = self._wsis_rulebook
        return torch.Size((rb.K, 2, rb.n_in))

    def to(self, *_a, **_k):
        return self


class lazy_pairs(object):
    def __enter__(self):
        global _LAZY_PAIRS
        self._old, _LAZY_PAIRS = _LAZY_PAIRS, True

    def __exit__(self, *exc):
        global _LAZY_PAIRS
        _LAZY_PAIRS = self._old
        return False


def _rulebook_of(indice_pairs, indice_pair_num, n_in, n_out, subm):
    rb = getattr(indice_pairs, "_wsis_rulebook", None)
    if rb is None:
        rb = W.rulebook_from_pairs(indice_pairs, indice_pair_num, n_in, n_out, subm)
        indice_pairs._wsis_rulebook = rb
    return rb


def _w3(filters):
    return filters.reshape(-1, filters.shape[-2], filters.shape[-1])


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse=False, subm=False,
                _prologue=None, _residual=None, _packed=None):
    """Forward of SubMConv / SparseConv / SparseInverseConv: indiceConv<T>, spconv_ops.h:253-349."""
    if filters.dtype != torch.float32:
        raise NotImplementedError
    n_feat = features.shape[0]
    if inverse:
        # features live on the coarse (couple's output) side, result on the fine (couple's input) side
        rb = _rulebook_of(indice_pairs, indice_pair_num, num_activate_out, n_feat, False)
        map_, flip, tiles = rb.nbr_in, 0, rb.tiles_in
    else:
        rb = _rulebook_of(indice_pairs, indice_pair_num, n_feat, num_activate_out, subm)
        (map_, flip), tiles = rb.fwd_map(), rb.tiles_out
    w3 = _w3(filters)
    use_tiles = W.get_precision() != "simt" and w3.shape[0] <= 32 and W.umma_supported(w3.shape[1], w3.shape[2])
    return W.sparse_conv(features, w3, map_, num_activate_out, flip, False, _prologue, _residual, _packed,
                         tiles=tiles() if use_tiles else None)


def indice_conv_backward(features, filters, out_bp, indice_pairs, indice_pair_num, inverse=False, subm=False,
                         _need_input_grad=True):
    """-> (input_bp, filters_bp): indiceConvBackward<T>, spconv_ops.h:351-433.  `_need_input_grad=False` (autograd knows
    that the features are a leaf without grad: the network's input conv) skips the data gradient and returns None."""
    if filters.dtype != torch.float32:
        raise NotImplementedError
    w3 = _w3(filters)
    K, Cin, Cout = w3.shape
    n_feat, n_out = features.shape[0], out_bp.shape[0]
    # dgrad contracts with W[k]^T: Cin_eff = Cout, Cout_eff = Cin
    use_tiles = W.get_precision() != "simt" and K <= 32 and W.umma_supported(Cout, Cin)
    if inverse:
        rb = _rulebook_of(indice_pairs, indice_pair_num, n_out, n_feat, False)
        din = W.sparse_conv(out_bp, w3, rb.nbr_out, n_feat, 0, True,
                            tiles=rb.tiles_out() if use_tiles else None) if _need_input_grad else None
        dw = W.sparse_conv_wgrad(features, rb.nbr_in, n_out, 0, out_bp, K, Cin, Cout, order=rb.order_hint("in"))
    else:
        rb = _rulebook_of(indice_pairs, indice_pair_num, n_feat, n_out, subm)
        fmap, fflip = rb.fwd_map()
        din = W.sparse_conv(out_bp, w3, rb.nbr_in, n_feat, 0, True,
                            tiles=rb.tiles_in() if use_tiles else None) if _need_input_grad else None
        dw = W.sparse_conv_wgrad(features, fmap, n_out, fflip, out_bp, K, Cin, Cout, order=rb.order_hint("out"))
    return din, dw.view(filters.shape)
