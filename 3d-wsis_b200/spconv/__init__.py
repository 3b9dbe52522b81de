"""Drop-in replacement for the reference's `spconv` package (spconv 1.0, llijiang fork):
modules/lib/spconv/spconv/__init__.py:20-96.  Same public names, same attribute surface, same parameter
layout; everything underneath runs on libwsis_b200.so (hash-table rulebooks, output-stationary tcgen05 conv).
"""
import numpy as np
import torch

from spconv import utils  # noqa: F401
from spconv.modules import SparseModule, SparseSequential
from spconv.conv import (SparseConv2d, SparseConv3d, SparseConvTranspose2d, SparseConvTranspose3d,  # noqa: F401
                         SparseInverseConv2d, SparseInverseConv3d, SubMConv2d, SubMConv3d, SparseConvolution)
from spconv import ops, functional  # noqa: F401


def scatter_nd(indices, updates, shape):
    """Dense scatter used by SparseConvTensor.dense() (reference __init__.py:29-42): no repeated indices."""
    out = torch.zeros(*shape, dtype=updates.dtype, device=updates.device)
    nd = indices.shape[-1]
    flat = indices.view(-1, nd)
    idx = tuple(flat[:, i] for i in range(nd)) + (Ellipsis,)
    out[idx] = updates.view(*(list(indices.shape[:-1]) + list(shape[nd:])))
    return out


class SparseConvTensor(object):
    """Container contract of reference __init__.py:44-83: features [N,C], indices int32 [N,4]=(b,x,y,z),
    spatial_shape, batch_size, indice_dict (rulebook cache keyed by indice_key), grid (unused: the hash-table
    rulebook needs no pre-allocated dense grid)."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        shape = [self.batch_size] + list(self.spatial_shape) + [self.features.shape[1]]
        res = scatter_nd(self.indices.long(), self.features, shape)
        if not channels_first:
            return res
        nd = len(self.spatial_shape)
        perm = list(range(0, nd + 1))
        perm.insert(1, nd + 1)
        return res.permute(*perm).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()


class RemoveGrid(SparseModule):
    def forward(self, x):
        x.grid = None
        return x
