"""Tensor-level host wrappers over the C ABI (include/wsis_b200.h).

PyTorch is only the allocator / stream provider here: every function takes torch tensors, allocates outputs
and workspaces with torch, and hands raw device pointers + the current CUDA stream to the library.  Nothing
in this file computes on the CPU; a CPU tensor where a CUDA tensor is required raises.
"""
import ctypes
import math
import os
import weakref

import torch

from ._lib import lib

# ---------------------------------------------------------------------------------------------------------
# configuration
# ---------------------------------------------------------------------------------------------------------
# "fp32": tcgen05 with bf16x3 split operands (fp32 contract, 1e-4); "bf16": tcgen05 with bf16 operands (1e-2);
# "simt": exact fp32 FFMA kernel everywhere.  Layers the tensor-core kernel cannot take (Cin % 32 != 0, e.g. the
# 6->32 input conv) always run on the SIMT kernel.
_PRECISION = os.environ.get("WSIS_PRECISION", "fp32")


# inference ECC-GRU: regenerate the edge filters on the tensor cores inside every step (csrc/ecc_umma.cu) instead of
# materialising [E,1024] with library GEMMs and streaming it 7 times; False = the streaming kernel (csrc/ecc.cu)
ECC_FUSED_FILTERS = os.environ.get("WSIS_ECC_FUSED", "1") != "0"


def set_precision(p):
    global _PRECISION
    assert p in ("fp32", "bf16", "simt"), p
    _PRECISION = p


def get_precision():
    return _PRECISION


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _i3(v):
    if not isinstance(v, (list, tuple)):
        v = [v] * 3
    v = [int(x) for x in v]
    assert len(v) == 3, "only 3-D sparse convolutions are supported"
    return v


def _c3(v):
    return (ctypes.c_int32 * 3)(*v)


def _cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("wsis_b200: %s must be a CUDA tensor (there is no CPU path for this op)" % name)
    return t


def _bytes(n, device):
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=device)


# Derived device-side images of parameters (pre-swizzled conv weights, folded BatchNorm scale/shift, packed ECC-GRU
# matrices) are cached per parameter (data_ptr, tensor._version).  In-place edits through `.data`
# (`w.data.copy_()`, `bn.weight.data.fill_()`: backbone_3D_WSIS.py:156-157 set_bn_init, spconv test_conv.py:360-361)
# do NOT bump `_version`, so the caches also carry this epoch: it is bumped by load_state_dict / .to() / .train() of
# the spconv modules and of the mirror Network, and callers that mutate `.data` after the first forward must call
# invalidate_caches() themselves (documented in INTEGRATION.md).
_CACHE_EPOCH = 0


def invalidate_caches():
    global _CACHE_EPOCH
    _CACHE_EPOCH += 1


def cache_epoch():
    return _CACHE_EPOCH


def launch_count():
    return int(lib().value("wsis_launch_count"))


# ---------------------------------------------------------------------------------------------------------
# rulebook
# ---------------------------------------------------------------------------------------------------------
def spatial_order(coords, spatial_shape, batch_size):
    """Morton-curve order of a coordinate set int32[N,4] -> int32[tile_pad(N)] (row ids, -1 padded).  Cached on
    the coords tensor: every rulebook of a level shares it."""
    cached = getattr(coords, "_wsis_order", None)
    if cached is not None:
        return cached
    N = coords.shape[0]
    dev = coords.device
    n_pad = lib().value("wsis_tile_pad", N)
    order = torch.empty((max(n_pad, 1),), dtype=torch.int32, device=dev)
    if N > 0:
        ws = _bytes(lib().value("wsis_spatial_order_ws_bytes", N), dev)
        lib().call("wsis_spatial_order", _ptr(coords), N, _c3(_i3(list(spatial_shape))), int(batch_size), _ptr(order),
                   _ptr(ws), _stream())
    try:
        coords._wsis_order = order
    except Exception:  # noqa: BLE001  (tensor subclasses that refuse attributes)
        pass
    return order


def identity_order(n, device):
    n_pad = lib().value("wsis_tile_pad", n)
    order = torch.empty((max(n_pad, 1),), dtype=torch.int32, device=device)
    if n > 0:
        lib().call("wsis_identity_order", n, _ptr(order), _stream())
    return order


class TileMap:
    """A neighbour map in the form conv_umma consumes: destination rows in tiles of 128 along `order`, one
    record per tile (include/wsis_b200.h, wsis_tile_records)."""

    def __init__(self, map_, n_dst, flip, order, identity=False):
        if identity:        # K = 1, map[r] = r: records written directly by wsis_tile_records_identity
            dev = order
            self.K, self.n_dst = 1, n_dst
            self.num_tiles = lib().value("wsis_tile_pad", n_dst) // 128
            self.stride = lib().value("wsis_tile_record_stride", 1)
            self.ustride = lib().value("wsis_tile_unique_stride", 1)
            self.order = torch.empty((max(self.num_tiles * 128, 1),), dtype=torch.int32, device=dev)
            self.records = _bytes(self.num_tiles * self.stride, dev)
            self.uidx = torch.empty((max(self.num_tiles * self.ustride, 4),), dtype=torch.int32, device=dev)
            self.meta = torch.empty((max(self.num_tiles, 1), 4), dtype=torch.int32, device=dev)
            self.stats = torch.empty((2,), dtype=torch.int32, device=dev)
            if n_dst > 0:
                lib().call("wsis_tile_records_identity", n_dst, _ptr(self.records), _ptr(self.uidx), _ptr(self.meta),
                           _ptr(self.stats), _ptr(self.order), _stream())
            return
        K = map_.shape[1]
        self.K, self.n_dst = K, n_dst
        self.order = order
        self.num_tiles = lib().value("wsis_tile_pad", n_dst) // 128
        self.stride = lib().value("wsis_tile_record_stride", K)
        self.ustride = lib().value("wsis_tile_unique_stride", K)
        # fixed-stride records: no size scan, no host sync
        self.records = _bytes(self.num_tiles * self.stride, map_.device)
        self.uidx = torch.empty((max(self.num_tiles * self.ustride, 4),), dtype=torch.int32, device=map_.device)
        self.meta = torch.empty((max(self.num_tiles, 1), 4), dtype=torch.int32, device=map_.device)
        self.stats = torch.zeros((2,), dtype=torch.int32, device=map_.device)
        if n_dst > 0:
            lib().call("wsis_tile_records", _ptr(map_), n_dst, K, int(flip), _ptr(order), _ptr(self.records),
                       _ptr(self.uidx), _ptr(self.meta), _ptr(self.stats), _stream())


def _tiles_of(map_, n_dst, flip):
    """Identity-order TileMap for callers that bring a bare neighbour map (cached on the map tensor)."""
    cache = getattr(map_, "_wsis_tiles", None)
    if cache is None:
        cache = {}
        map_._wsis_tiles = cache
    if flip not in cache:
        cache[flip] = TileMap(map_, n_dst, flip, identity_order(n_dst, map_.device))
    return cache[flip]


class Rulebook:
    """Output-stationary rulebook: nbr_in[i,k] = out row, nbr_out[o,k] = in row (see include/wsis_b200.h).

    Replaces the (indice_pairs, indice_pair_num) pair of spconv_ops.h:27-137; `pairs()` still materialises the
    reference-format tensors (conv.py:152 stores them in indice_dict).  `tiles_out()` / `tiles_in()` are the
    spatially tiled maps the tensor-core kernel walks (destination = output side / input side)."""

    def __init__(self, kind, K, n_in, n_out, nbr_in, nbr_out, out_coords, in_coords=None, in_shape=None,
                 out_shape=None, batch_size=1):
        self.kind, self.K, self.n_in, self.n_out = kind, K, n_in, n_out
        self.nbr_in, self.nbr_out, self.out_coords = nbr_in, nbr_out, out_coords
        self.in_coords, self.in_shape, self.out_shape, self.batch_size = in_coords, in_shape, out_shape, batch_size
        self._pairs = None
        self._tiles = {}

    def pairs(self):
        # The reference-format tensors carry a strong reference to this rulebook (`pairs._wsis_rulebook`, so that the
        # rulebook lives as long as indice_dict holds the pairs); the way back is WEAK: a strong one would make every
        # rulebook of every step (pairs [K,2,N] = 117 MB at level 1, tile records, maps) a reference cycle that only the
        # cyclic collector frees -- in training that showed up as GBs of dead rulebooks and 100-380 ms allocator stalls.
        if self._pairs is not None:
            cached = tuple(r() for r in self._pairs)
            if all(t is not None for t in cached):
                return cached
        dev = self.nbr_in.device
        N, K = self.n_in, self.K
        pairs = torch.empty((K, 2, N), dtype=torch.int32, device=dev)
        num = torch.empty((K,), dtype=torch.int32, device=dev)
        if N > 0:
            pos = torch.empty((K * N + 1,), dtype=torch.int32, device=dev)
            ws = _bytes(lib().value("wsis_scan_ws_bytes", K * N), dev)
            lib().call("wsis_pairs_from_nbr", _ptr(self.nbr_in), N, K, _ptr(pairs), _ptr(num), _ptr(pos), _ptr(ws), _stream())
        else:
            num.zero_()
        self._pairs = [weakref.ref(pairs), weakref.ref(num)]
        return pairs, num

    # (map, flip) for: y[dst] = sum_k x[map[dst,k]] W[k]
    def fwd_map(self):
        # submanifold rulebooks of ODD kernels without dilation are symmetric (nbr_out[o,k] == nbr_in[o,K-1-k]); any
        # other submanifold rulebook carries its own output-side map (rulebook_subm builds it)
        return (self.nbr_in, 1) if (self.kind == "subm" and self.nbr_out is None) else (self.nbr_out, 0)

    def bwd_map(self):  # maps rows of the conv INPUT side to rows of the OUTPUT side
        return (self.nbr_in, 0)

    def _order(self, side):
        coords, shape = (self.out_coords, self.out_shape) if side == "out" else (self.in_coords, self.in_shape)
        n = self.n_out if side == "out" else self.n_in
        if coords is None or shape is None:
            return identity_order(n, self.nbr_in.device)
        return spatial_order(coords, shape, self.batch_size)

    def order_hint(self, side):
        """Morton order of one side's rows when the coordinates are known, else None (= natural order)."""
        coords, shape = (self.out_coords, self.out_shape) if side == "out" else (self.in_coords, self.in_shape)
        return None if coords is None or shape is None else spatial_order(coords, shape, self.batch_size)

    def tiles_out(self):
        """Destination = the conv's output rows (forward of subm / strided conv, dgrad of the inverse conv)."""
        if "out" not in self._tiles:
            m, flip = self.fwd_map()
            self._tiles["out"] = TileMap(m, self.n_out, flip, self._order("out"))
        return self._tiles["out"]

    def tiles_in(self):
        """Destination = the conv's input rows (dgrad of subm / strided conv, forward of the inverse conv)."""
        if "in" not in self._tiles:
            self._tiles["in"] = TileMap(self.nbr_in, self.n_in, 0, self._order("in"))
        return self._tiles["in"]


def rulebook_subm(indices, spatial_shape, ksize=3, dilation=1, batch_size=None):
    """Submanifold rulebook: replaces getIndicePair<3> with subM=1 (spconv_ops.h:86-102)."""
    indices = _cuda(indices, "indices")
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4
    indices = indices.contiguous()
    ks, dil, shape = _i3(ksize), _i3(dilation), _i3(list(spatial_shape))
    K = ks[0] * ks[1] * ks[2]
    N = indices.shape[0]
    dev = indices.device
    nbr = torch.empty((N, K), dtype=torch.int32, device=dev)
    if N > 0:
        slots = lib().value("wsis_hash_slots", N)
        keys = torch.empty((slots,), dtype=torch.int64, device=dev)
        vals = torch.empty((slots,), dtype=torch.int32, device=dev)
        lib().call("wsis_rulebook_subm", _ptr(indices), N, _c3(ks), _c3(dil), _c3(shape), _ptr(keys), _ptr(vals), slots,
                   _ptr(nbr), _stream())
    rb = Rulebook("subm", K, N, N, nbr, None, indices, indices, shape, shape, _batch_of(batch_size))
    if any(k % 2 == 0 for k in ks) or any(d != 1 for d in dil):
        # Like the reference (spconv_ops.h:74-77) the submanifold padding is ks/2 whatever the dilation, so for even
        # kernels or dilation > 1 the pair (k, i -> o) has no mirror image (K-1-k, o -> i): the forward / wgrad map
        # cannot be read off nbr_in by flipping the offset.  Build the output-side map from the pairs themselves.
        pairs, num = rb.pairs()
        nbr_out = torch.full((N, K), -1, dtype=torch.int32, device=dev)
        if N > 0:
            lib().call("wsis_nbr_from_pairs", _ptr(pairs), _ptr(num), pairs.shape[2], K, 1, _ptr(nbr_out), _stream())
        rb.nbr_out = nbr_out
    return rb


def _batch_of(batch_size):
    # the Morton key reserves ceil(log2(batch)) bits; without a hint allow for 65536 scenes per batch
    return 65536 if batch_size is None else max(int(batch_size), 1)


def conv_output_shape(spatial_shape, ksize, stride, padding, dilation):
    """ops.get_conv_output_size, spconv/ops.py:19-30."""
    s, k, st, p, d = _i3(list(spatial_shape)), _i3(ksize), _i3(stride), _i3(padding), _i3(dilation)
    return [(s[i] + 2 * p[i] - d[i] * (k[i] - 1) - 1) // st[i] + 1 for i in range(3)]


def rulebook_conv(indices, spatial_shape, ksize, stride, padding=0, dilation=1, batch_size=None):
    """Strided sparse-conv rulebook: replaces getIndicePair<3> with subM=0 (spconv_ops.h:103-136).
    One host sync to learn the output count (the reference syncs for the same reason, :126-135)."""
    indices = _cuda(indices, "indices")
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4
    indices = indices.contiguous()
    ks, st, pad, dil = _i3(ksize), _i3(stride), _i3(padding), _i3(dilation)
    oshape = conv_output_shape(spatial_shape, ks, st, pad, dil)
    if min(oshape) <= 0:
        raise RuntimeError("wsis_b200: empty output spatial shape %s" % (oshape,))
    K = ks[0] * ks[1] * ks[2]
    N = indices.shape[0]
    dev = indices.device
    nbr_in = torch.empty((N, K), dtype=torch.int32, device=dev)
    if N == 0:
        return Rulebook("conv", K, 0, 0, nbr_in, torch.empty((0, K), dtype=torch.int32, device=dev),
                        torch.empty((0, 4), dtype=torch.int32, device=dev)), oshape
    tpi = 1
    for d in range(3):
        tpi *= (ks[d] + st[d] - 1) // st[d]
    slots = lib().value("wsis_hash_slots", N * tpi)
    keys = torch.empty((slots,), dtype=torch.int64, device=dev)
    vals = torch.empty((slots,), dtype=torch.int32, device=dev)
    rank = torch.empty((N * K + 1,), dtype=torch.int32, device=dev)
    scan_ws = _bytes(lib().value("wsis_scan_ws_bytes", N * K), dev)
    n_out_dev = torch.empty((1,), dtype=torch.int32, device=dev)
    lib().call("wsis_rulebook_conv_count", _ptr(indices), N, _c3(ks), _c3(st), _c3(pad), _c3(dil), _c3(oshape),
               _ptr(keys), _ptr(vals), slots, _ptr(nbr_in), _ptr(rank), _ptr(scan_ws), _ptr(n_out_dev), _stream())
    n_out = int(n_out_dev.item())
    out_coords = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
    nbr_out = torch.empty((n_out, K), dtype=torch.int32, device=dev)
    slot_rank = torch.empty((slots,), dtype=torch.int32, device=dev)
    lib().call("wsis_rulebook_conv_fill", _ptr(indices), N, K, _ptr(keys), _ptr(vals), slots, _ptr(nbr_in), _ptr(rank),
               _ptr(slot_rank), n_out, _ptr(out_coords), _ptr(nbr_out), _stream())
    return Rulebook("conv", K, N, n_out, nbr_in, nbr_out, out_coords, indices, _i3(list(spatial_shape)), oshape,
                    _batch_of(batch_size)), oshape


def rulebook_from_pairs(pairs, num, n_in, n_out, subm):
    """For callers that bring their own reference-format pairs (ops.indice_conv drop-in, spconv/ops.py:101-116)."""
    pairs = _cuda(pairs, "indice_pairs").contiguous()
    num = _cuda(num, "indice_pair_num").contiguous()
    assert pairs.dtype == torch.int32 and num.dtype == torch.int32
    K, stride = pairs.shape[0], pairs.shape[2]
    dev = pairs.device
    nbr_in = torch.full((n_in, K), -1, dtype=torch.int32, device=dev)
    lib().call("wsis_nbr_from_pairs", _ptr(pairs), _ptr(num), stride, K, 0, _ptr(nbr_in), _stream())
    nbr_out = None
    if not subm:
        nbr_out = torch.full((n_out, K), -1, dtype=torch.int32, device=dev)
        lib().call("wsis_nbr_from_pairs", _ptr(pairs), _ptr(num), stride, K, 1, _ptr(nbr_out), _stream())
    rb = Rulebook("subm" if subm else "conv", K, n_in, n_out, nbr_in, nbr_out, None)
    rb._pairs = [weakref.ref(pairs), weakref.ref(num)]     # weak: the caller's pairs tensor points back at `rb`
    return rb


# ---------------------------------------------------------------------------------------------------------
# sparse convolution
# ---------------------------------------------------------------------------------------------------------
class PackedWeights:
    """Pre-swizzled bf16 (hi[,mid]) image of a conv weight for the tcgen05 kernel; cached per (weight version)."""

    def __init__(self):
        self.key = None
        self.buf = None

    def get(self, weight3, transpose_w, precision):
        Cin, Cout = (weight3.shape[2], weight3.shape[1]) if transpose_w else (weight3.shape[1], weight3.shape[2])
        key = (weight3.data_ptr(), weight3._version, tuple(weight3.shape), transpose_w, precision, _CACHE_EPOCH)
        if key != self.key:
            K = weight3.shape[0]
            nbytes = lib().value("wsis_conv_pack_bytes", K, Cin, Cout, precision)
            buf = _bytes(nbytes, weight3.device)
            lib().call("wsis_conv_pack_weights", _ptr(weight3), K, Cin, Cout, int(transpose_w), precision, _ptr(buf),
                       _stream())
            self.key, self.buf = key, buf
        return self.buf


def umma_supported(Cin, Cout):
    return bool(lib().value("wsis_conv_umma_supported", int(Cin), int(Cout)))


def sparse_conv(src, weight3, map_, n_dst, flip, transpose_w=False, prologue=None, residual=None, packed=None,
                precision=None, tiles=None):
    """dst[r] = residual[r] + sum_k prologue(src[map[r,k']]) @ W[k]  (W[k]^T when transpose_w).

    src f32[n_src, Cin]; weight3 f32[K, Cin_w, Cout_w]; map_ int32[n_dst, K]; `tiles` = the TileMap of
    (map_, flip) when the caller has one (Rulebook.tiles_out()/tiles_in()), else an identity-order one is built."""
    src = _cuda(src, "features")
    if src.dtype != torch.float32:
        raise RuntimeError("wsis_b200: features must be float32, got %s" % src.dtype)
    src = src.contiguous()
    weight3 = weight3.detach()
    if weight3.dtype != torch.float32 or not weight3.is_cuda:
        raise RuntimeError("wsis_b200: filters must be a float32 CUDA tensor, got %s on %s" % (weight3.dtype, weight3.device))
    if not weight3.is_contiguous():
        weight3 = weight3.contiguous()
    K = weight3.shape[0]
    Cin, Cout = (weight3.shape[2], weight3.shape[1]) if transpose_w else (weight3.shape[1], weight3.shape[2])
    if src.shape[1] != Cin:
        raise RuntimeError("wsis_b200: feature width %d != filter in-planes %d" % (src.shape[1], Cin))
    if map_.shape[0] != n_dst or map_.shape[1] != K:
        raise RuntimeError("wsis_b200: neighbour map shape %s does not match (n_dst=%d, K=%d)" %
                           (tuple(map_.shape), n_dst, K))
    dst = torch.empty((n_dst, Cout), dtype=torch.float32, device=src.device)
    if n_dst == 0:
        return dst
    scale = shift = None
    relu = 0
    if prologue is not None:
        scale, shift, relu = prologue
        scale, shift = scale.contiguous(), shift.contiguous()
    if residual is not None:
        residual = residual.contiguous()
        assert residual.shape == dst.shape and residual.dtype == torch.float32
    precision = precision or _PRECISION
    if precision != "simt" and K <= 32 and umma_supported(Cin, Cout):
        prec = 3 if precision == "fp32" else 1
        packed = packed if packed is not None else PackedWeights()
        buf = packed.get(weight3, bool(transpose_w), prec)
        tiles = tiles if tiles is not None else _tiles_of(map_, n_dst, int(flip))
        assert tiles.n_dst == n_dst and tiles.K == K
        lib().call("wsis_conv_umma", _ptr(src), _ptr(tiles.records), _ptr(tiles.uidx), _ptr(tiles.meta), _ptr(tiles.order),
                   tiles.num_tiles, K, _ptr(buf), Cin, Cout, prec, _ptr(scale), _ptr(shift), int(relu), _ptr(residual),
                   _ptr(dst), _stream())
    else:
        lib().call("wsis_conv_simt", _ptr(src), _ptr(map_), n_dst, K, int(flip), _ptr(weight3), int(transpose_w), Cin,
                   Cout, _ptr(scale), _ptr(shift), int(relu), _ptr(residual), _ptr(dst), _stream())
    return dst


def _identity_tiles(n, device, holder=None):
    """(map int32[n,1] = iota, TileMap) of the identity rulebook: what a 1x1 convolution / a Linear over rows looks
    like to conv_umma.  Cached on `holder` (the coordinate tensor of the level) when given."""
    cached = getattr(holder, "_wsis_ident", None) if holder is not None else None
    if cached is not None and cached[0].shape[0] == n:
        return cached
    tiles = TileMap(None, n, 0, device, identity=True)
    map_ = tiles.order[:n].view(n, 1)                     # the identity order doubles as the [n, 1] neighbour map
    cached = (map_, tiles)
    if holder is not None:
        try:
            holder._wsis_ident = cached
        except Exception:  # noqa: BLE001
            pass
    return cached


def dense_rows(src, weight, packed=None, holder=None, precision=None, linear_layout=False):
    """src f32[n, Cin] @ W on the tensor-core conv kernel (a 1x1 submanifold conv is a sparse conv with the identity
    rulebook: conv.py:113-119 does it with torch.mm).  weight f32[Cin, Cout], or the nn.Linear layout f32[Cout, Cin]
    with linear_layout=True (y = x @ weight^T).  Returns None when the shape is not supported (the caller keeps its
    library GEMM)."""
    Cin, Cout = (weight.shape[1], weight.shape[0]) if linear_layout else weight.shape
    precision = precision or _PRECISION
    if precision == "simt" or not src.is_cuda or src.dtype != torch.float32 or src.shape[0] == 0 \
            or not umma_supported(Cin, Cout):
        return None
    n = src.shape[0]
    map_, tiles = _identity_tiles(n, src.device, holder)
    return sparse_conv(src, weight.unsqueeze(0), map_, n, 0, bool(linear_layout), packed=packed, precision=precision,
                       tiles=tiles)


def sparse_conv_wgrad(src, map_, n_dst, flip, grad_out, K, Cin, Cout, prologue=None, order=None, precision=None):
    """dW[k] = sum_r prologue(src[map[r,k']])^T grad_out[r]   -> f32[K, Cin, Cout].
    Tensor-core kernel (csrc/wgrad_umma.cu) when the shape allows it and the precision is not "simt"; `order` = the
    Morton order of the destination rows (TileMap.order) when the caller has one."""
    src = _cuda(src, "features").contiguous()
    grad_out = grad_out.contiguous()
    dW = torch.empty((K, Cin, Cout), dtype=torch.float32, device=src.device)
    scale = shift = None
    relu = 0
    if prologue is not None:
        scale, shift, relu = prologue
        scale, shift = scale.contiguous(), shift.contiguous()
    precision = precision or _PRECISION
    if precision != "simt" and Cin < 32 and n_dst > 0 and lib().value("wsis_conv_wgrad_umma_supported", K, 32, Cout):
        # the network's input conv (6 -> 32): zero-pad the rows to one 32-channel block and slice the result
        pad = torch.zeros((src.shape[0], 32), dtype=torch.float32, device=src.device)
        pad[:, :Cin] = src
        ps = None
        if prologue is not None:
            ps = (torch.nn.functional.pad(scale, (0, 32 - Cin)), torch.nn.functional.pad(shift, (0, 32 - Cin)), relu)
        return sparse_conv_wgrad(pad, map_, n_dst, flip, grad_out, K, 32, Cout, ps, order, precision)[:, :Cin, :].contiguous()
    if precision != "simt" and lib().value("wsis_conv_wgrad_umma_supported", K, Cin, Cout):
        lib().call("wsis_conv_wgrad_umma", _ptr(src), _ptr(map_), _ptr(order), n_dst, K, int(flip), _ptr(grad_out), Cin,
                   Cout, _ptr(scale), _ptr(shift), int(relu), 3 if precision == "fp32" else 1, _ptr(dW), _stream())
        return dW
    lib().call("wsis_conv_wgrad", _ptr(src), _ptr(map_), n_dst, K, int(flip), _ptr(grad_out), Cin, Cout, _ptr(scale),
               _ptr(shift), int(relu), _ptr(dW), _stream())
    return dW


def affine_relu(x, scale, shift, relu=True, out=None):
    """Eval-mode BatchNorm1d folded to y = x*scale+shift, optionally ReLU; one pass."""
    x = _cuda(x, "x").contiguous()
    out = torch.empty_like(x) if out is None else out
    scale, shift = scale.contiguous(), shift.contiguous()  # temporaries must outlive the launch call
    lib().call("wsis_affine_relu", _ptr(x), x.shape[0], x.shape[1], _ptr(scale), _ptr(shift), int(relu), _ptr(out),
               _stream())
    return out


# ---------------------------------------------------------------------------------------------------------
# voxelization (pointgroup_ops contract)
# ---------------------------------------------------------------------------------------------------------
def voxelization_idx(coords, batch_size=None, mode=4):
    """coords int64[N,4] -> (voxel_locs int64[M,4], p2v int32[N], v2p int32[M,1+maxActive]).
    CPU tensors use the CUDA-free host routine (DataLoader workers, scannetv2_dataset.py:449); CUDA tensors
    use the device hash-table path.  Both number voxels in first-occurrence order."""
    assert coords.dtype == torch.int64 and coords.dim() == 2 and coords.shape[1] == 4
    coords = coords.contiguous()
    N = coords.shape[0]
    if not coords.is_cuda:
        ma = ctypes.c_int32(0)
        p2v = torch.empty((N,), dtype=torch.int32)
        # phase 1 hashes the points once (voxel count, max points per voxel, p2v); phase 2 is a linear fill
        M = lib().value("wsis_voxelize_idx_host", _ptr(coords), N, None, _ptr(p2v), None, 0, ctypes.byref(ma))
        if M < 0:
            raise RuntimeError("voxelization_idx: " + lib().last_error())
        locs = torch.empty((M, 4), dtype=torch.int64)
        v2p = torch.empty((M, 1 + ma.value), dtype=torch.int32)
        lib().call("wsis_voxelize_idx_host_fill", _ptr(coords), N, _ptr(p2v), M, _ptr(locs), _ptr(v2p), 1 + ma.value)
        return locs, p2v, v2p
    dev = coords.device
    p2v = torch.empty((N,), dtype=torch.int32, device=dev)
    counts = torch.empty((3,), dtype=torch.int32, device=dev)
    ws = _bytes(lib().value("wsis_voxelize_ws_bytes", N), dev)
    lib().call("wsis_voxelize_idx_count", _ptr(coords), N, _ptr(p2v), _ptr(counts), _ptr(ws), _stream())
    M, ma, err = (int(x) for x in counts.tolist())
    if err:
        raise RuntimeError("voxelization_idx: coordinates must lie in [0, 65535]")
    locs = torch.empty((M, 4), dtype=torch.int64, device=dev)
    v2p = torch.empty((M, 1 + ma), dtype=torch.int32, device=dev)
    lib().call("wsis_voxelize_idx_fill", _ptr(coords), N, _ptr(p2v), M, ma, _ptr(locs), _ptr(v2p), _ptr(ws), _stream())
    return locs, p2v, v2p


def voxelization_fwd(feats, v2p):
    feats = _cuda(feats, "feats").contiguous()
    v2p = _cuda(v2p, "v2p_map").contiguous()
    assert feats.dtype == torch.float32 and v2p.dtype == torch.int32
    M, C = v2p.shape[0], feats.shape[1]
    out = torch.empty((M, C), dtype=torch.float32, device=feats.device)
    lib().call("wsis_voxelize_mean_fwd", _ptr(feats), _ptr(v2p), M, v2p.shape[1], C, _ptr(out), _stream())
    return out


def voxelization_bwd(dout, v2p, n_points):
    dout = _cuda(dout, "d_output_feats").contiguous()
    v2p = v2p.contiguous()
    M, C = dout.shape
    df = torch.empty((n_points, C), dtype=torch.float32, device=dout.device)
    lib().call("wsis_voxelize_mean_bwd", _ptr(dout), _ptr(v2p), M, v2p.shape[1], C, n_points, _ptr(df), _stream())
    return df


# ---------------------------------------------------------------------------------------------------------
# segmented reductions (torch_scatter.scatter contract)
# ---------------------------------------------------------------------------------------------------------
class SegmentIndex:
    """CSR of an unsorted id vector, built once per scene and shared by every reduction over it."""

    def __init__(self, ids, num_segments):
        ids = _cuda(ids, "index")
        assert ids.dtype == torch.int64 and ids.dim() == 1
        ids = ids.contiguous()
        self.ids = ids
        self.n, self.S = ids.shape[0], int(num_segments)
        dev = ids.device
        self.order = torch.empty((max(self.n, 1),), dtype=torch.int32, device=dev)
        self.offsets = torch.empty((self.S + 1,), dtype=torch.int32, device=dev)
        ws = _bytes(lib().value("wsis_segment_csr_ws_bytes", self.n, self.S), dev)
        lib().call("wsis_segment_csr", _ptr(ids), self.n, self.S, _ptr(self.order), _ptr(self.offsets), _ptr(ws),
                   _stream())

    def validate(self):
        """Raises if an id lies outside [0, S): the CSR kernel clamps such ids into the first / last segment instead of
        corrupting memory, which would silently merge foreign rows into that segment's reduction (a wrong
        `num_superpoints` in a batch dict).  One host sync: for set-up code and the torch_scatter drop-in, not for the
        per-step path."""
        if self.n:
            lo, hi = int(self.ids.min().item()), int(self.ids.max().item())
            if lo < 0 or hi >= self.S:
                raise RuntimeError("wsis_b200: segment ids span [%d, %d] but num_segments is %d" % (lo, hi, self.S))
        return self


_REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2}


def segment_reduce(src, seg, reduce="mean", gather=None):
    """out[s] = reduce over rows i with ids[i]==s of src[gather[i] if gather is given else i]."""
    src = _cuda(src, "src")
    squeeze = src.dim() == 1
    src2 = (src.unsqueeze(1) if squeeze else src).contiguous()
    assert src2.dtype == torch.float32
    out = torch.empty((seg.S, src2.shape[1]), dtype=torch.float32, device=src.device)
    if gather is not None:
        assert gather.dtype == torch.int32
        gather = gather.contiguous()
    lib().call("wsis_segment_reduce", _ptr(src2), _ptr(gather), _ptr(seg.order), _ptr(seg.offsets), seg.S, src2.shape[1],
               _REDUCE[reduce], _ptr(out), _stream())
    return out[:, 0] if squeeze else out


def scatter(src, index, dim=0, reduce="sum", dim_size=None):
    """torch_scatter.scatter(src, index, dim=0, reduce=...) drop-in for the calls at backbone_3D_WSIS.py:188,225,
    232,244.  Like torch_scatter, dim_size=None costs one host sync for index.max()."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
        if index.numel() > 0 and int(index.min().item()) < 0:
            raise RuntimeError("wsis_b200: scatter index must be non-negative")
        return segment_reduce(src, SegmentIndex(index, dim_size), reduce)
    return segment_reduce(src, SegmentIndex(index, dim_size).validate(), reduce)    # like torch_scatter: bad ids raise


def gather_rows(src, idx):
    src = _cuda(src, "src").contiguous()
    idx = idx.contiguous()
    assert idx.dtype == torch.int32 and src.dtype == torch.float32
    out = torch.empty((idx.shape[0], src.shape[1]), dtype=torch.float32, device=src.device)
    lib().call("wsis_gather_rows", _ptr(src), _ptr(idx), idx.shape[0], src.shape[1], _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------------------
# affinity + random walk
# ---------------------------------------------------------------------------------------------------------
def edge_attention(q, k, v, ecc, centers, edge_u, edge_v, eseg, pos_mlp):
    """backbone_3D_WSIS.py:209-249 in one kernel.  Returns (edge_affinity f32[E], sp_feat f32[S,64])."""
    q, k, v, ecc, centers = (_cuda(t, "attention input").contiguous() for t in (q, k, v, ecc, centers))
    S, D = q.shape
    E = edge_u.shape[0]
    aff = torch.empty((E,), dtype=torch.float32, device=q.device)
    sp = torch.empty((S, D), dtype=torch.float32, device=q.device)
    edge_u, edge_v, pos_mlp = edge_u.contiguous(), edge_v.contiguous(), pos_mlp.contiguous()
    assert edge_u.dtype == torch.int64 and edge_v.dtype == torch.int64 and pos_mlp.numel() == 81
    lib().call("wsis_edge_attention", _ptr(q), _ptr(k), _ptr(v), _ptr(ecc), _ptr(centers), _ptr(edge_u), _ptr(edge_v),
               _ptr(eseg.order), _ptr(eseg.offsets), S, E, D, _ptr(pos_mlp), _ptr(aff), _ptr(sp), _stream())
    return aff, sp


def _head_pack(head):
    """Sequential(Linear, BatchNorm1d, ReLU, Linear) (backbone_3D_WSIS.py:57-62) -> the folded, transposed parameter
    images wsis_mlp_head reads; cached on the module per parameter version / cache epoch."""
    l1, bn, _, l2 = head[0], head[1], head[2], head[3]
    ps = [l1.weight, l1.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, l2.weight, l2.bias]
    key = tuple((p._version, p.data_ptr()) for p in ps if p is not None) + (_CACHE_EPOCH,)
    cache = getattr(head, "_wsis_head", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            inv = torch.rsqrt(bn.running_var.float() + bn.eps)
            scale = (bn.weight.float() if bn.weight is not None else torch.ones_like(inv)) * inv
            shift = (bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)) - bn.running_mean.float() * scale
            w1t = (l1.weight.float() * scale.unsqueeze(1)).t().contiguous()             # [Cin, H]
            b1 = l1.bias.float() if l1.bias is not None else torch.zeros_like(scale)
            t1 = (b1 * scale + shift).contiguous()
            cout = l2.weight.shape[0]
            cp = lib().value("wsis_mlp_head_coutp", cout)
            w2t = torch.zeros((l2.weight.shape[1], cp), dtype=torch.float32, device=l2.weight.device)
            w2t[:, :cout] = l2.weight.float().t()
            b2 = torch.zeros((cp,), dtype=torch.float32, device=l2.weight.device)
            if l2.bias is not None:
                b2[:cout] = l2.bias.float()
        cache = (key, w1t, t1, w2t, b2, cout)
        head._wsis_head = cache
    return cache[1:]


def mlp_head_supported(head, x):
    """Linear(C,C) -> BatchNorm1d(eval, running stats) -> ReLU -> Linear(C, <=32) with C in {32, 64} on CUDA fp32 rows."""
    import torch.nn as nn
    mods = list(head._modules.values()) if hasattr(head, "_modules") else []
    if len(mods) != 4 or not (isinstance(mods[0], nn.Linear) and isinstance(mods[1], nn.BatchNorm1d)
                              and type(mods[2]) is nn.ReLU and isinstance(mods[3], nn.Linear)):
        return False
    l1, bn, _, l2 = mods
    C = l1.in_features
    return (C in (32, 64) and l1.out_features == C and l2.in_features == C and l2.out_features <= 32
            and not bn.training and bn.track_running_stats and x.is_cuda and x.dtype == torch.float32
            and x.dim() == 2 and x.shape[1] == C)


def mlp_head(head, x, gather=None):
    """head(x[gather]) in one kernel (csrc/heads.cu); `gather` int32[n] or None."""
    w1t, t1, w2t, b2, cout = _head_pack(head)
    x = _cuda(x, "head input").contiguous()
    n = x.shape[0] if gather is None else gather.shape[0]
    out = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    if gather is not None:
        assert gather.dtype == torch.int32
        gather = gather.contiguous()
    lib().call("wsis_mlp_head", _ptr(x), _ptr(gather), n, x.shape[1], w1t.shape[1], cout, _ptr(w1t), _ptr(t1), _ptr(w2t),
               _ptr(b2), _ptr(out), _stream())
    return out


def pack_ecc_gru(cell):
    """GRUCellEx parameters -> float[wsis_ecc_gru_param_floats()] in the kernel's (transposed) layout."""
    ig = cell._modules["ig"]
    buf = torch.cat([ig.weight.detach().t().reshape(-1), ig.bias.detach().reshape(-1),
                     cell.weight_ih.detach().t().reshape(-1), cell.weight_hh.detach().t().reshape(-1),
                     cell.bias_ih.detach().reshape(-1), cell.bias_hh.detach().reshape(-1)]).float().contiguous()
    assert buf.numel() == lib().value("wsis_ecc_gru_param_floats")
    return buf


def ecc_gru(hx, filters, src, tseg, params, nrepeats, layernorm=True, eps=1e-5, cat_all=True):
    """`nrepeats` steps of the edge-conditioned GRU (spg_modules.py:152-185, 226-253), one kernel per step.
    hx f32[S,32]; filters f32[E,1024]; src int64[E]; tseg = SegmentIndex of the edge TARGETS."""
    hx = _cuda(hx, "hx").contiguous()
    filters = _cuda(filters, "filters").contiguous()
    src = src.contiguous()
    S, F = hx.shape
    if F != 32 or hx.dtype != torch.float32 or filters.dtype != torch.float32 or filters.shape[1] != F * F:
        raise RuntimeError("wsis_b200: ecc_gru needs float32 nfeat=32 states and [E,1024] filters")
    assert src.dtype == torch.int64 and tseg.S == S and tseg.n == src.shape[0] == filters.shape[0]
    width = F * (nrepeats + 1)
    cat = torch.empty((S, width), dtype=torch.float32, device=hx.device) if cat_all else None
    if cat_all:
        cat[:, :F] = hx
    h = hx
    for r in range(nrepeats):
        out = torch.empty_like(h)
        cat_ptr = ctypes.c_void_p(cat.data_ptr() + 4 * F * (r + 1)) if cat_all else None
        lib().call("wsis_ecc_gru_step", _ptr(h), _ptr(filters), _ptr(src), _ptr(tseg.order), _ptr(tseg.offsets), S,
                   _ptr(params), int(layernorm), float(eps), _ptr(out), cat_ptr, width, _stream())
        h = out
    return cat if cat_all else h


def ecc_fnet_supported(fnet):
    """Filter network of the 3D-WSIS ECC layer (graphnet.py:21-39 with widths 13-32-128-64-1024, BatchNorm after the
    third Linear, eval mode): Linear, ReLU, Linear, ReLU, Linear, BatchNorm1d, ReLU, Linear."""
    import torch.nn as nn
    m = list(fnet._modules.values()) if hasattr(fnet, "_modules") else []
    kinds = [nn.Linear, nn.ReLU, nn.Linear, nn.ReLU, nn.Linear, nn.BatchNorm1d, nn.ReLU, nn.Linear]
    if len(m) != len(kinds) or not all(isinstance(a, k) for a, k in zip(m, kinds)):
        return False
    shapes = [(m[0].in_features, m[0].out_features), (m[2].in_features, m[2].out_features),
              (m[4].in_features, m[4].out_features), (m[7].in_features, m[7].out_features)]
    return shapes == [(13, 32), (32, 128), (128, 64), (64, 1024)] and not m[5].training and m[5].track_running_stats


def pack_ecc_fnet(fnet):
    """-> (edge-MLP parameter pack f32[wsis_ecc_edge_mlp_param_floats()], W4 operand image, b4 f32[1024] or None);
    cached on the module per parameter version / cache epoch."""
    m = list(fnet._modules.values())
    ps = [p for mod in m for p in list(mod.parameters()) + list(mod.buffers())]
    key = tuple((p._version, p.data_ptr()) for p in ps) + (_CACHE_EPOCH,)
    cache = getattr(fnet, "_wsis_fnet", None)
    if cache is None or cache[0] != key:
        l1, l2, l3, bn, l4 = m[0], m[2], m[4], m[5], m[7]
        with torch.no_grad():
            inv = torch.rsqrt(bn.running_var.float() + bn.eps)
            scale = (bn.weight.float() if bn.weight is not None else torch.ones_like(inv)) * inv
            shift = (bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)) - bn.running_mean.float() * scale

            def bias(l):
                return l.bias.float() if l.bias is not None else torch.zeros(l.out_features, device=l.weight.device)
            params = torch.cat([l1.weight.float().t().reshape(-1), bias(l1), l2.weight.float().t().reshape(-1), bias(l2),
                                (l3.weight.float() * scale.unsqueeze(1)).t().reshape(-1), bias(l3) * scale + shift]).contiguous()
            assert params.numel() == lib().value("wsis_ecc_edge_mlp_param_floats")
            w4 = l4.weight.detach().float().contiguous()
            w4p = _bytes(lib().value("wsis_ecc_w4_bytes"), w4.device)
            lib().call("wsis_ecc_pack_w4", _ptr(w4), _ptr(w4p), _stream())
            b4 = l4.bias.detach().float().contiguous() if l4.bias is not None else None
        cache = (key, params, w4p, b4)
        fnet._wsis_fnet = cache
    return cache[1:]


def ecc_gru_fused(hx, fnet, edgefeats, src, tseg, params, nrepeats, layernorm=True, eps=1e-5, cat_all=True,
                  edges_sorted=False):
    """`nrepeats` ECC-GRU steps WITHOUT materialised edge filters (csrc/ecc_umma.cu): the filter network's hidden
    layers run once per edge, the 64 -> 1024 filter layer is regenerated on the tensor cores inside every step.
    hx f32[S,32]; edgefeats f32[E,13]; src int64[E]; tseg = SegmentIndex of the edge TARGETS; `edges_sorted`: the
    edge arrays are already in target order (ecc/GraphConvInfo.py:50-76), so tseg.order is the identity."""
    hx = _cuda(hx, "hx").contiguous()
    edgefeats = _cuda(edgefeats, "edge features").contiguous()
    src = src.contiguous()
    S, F = hx.shape
    E = src.shape[0]
    assert F == 32 and hx.dtype == torch.float32 and edgefeats.shape == (E, 13) and tseg.S == S and tseg.n == E
    mlp, w4p, b4 = pack_ecc_fnet(fnet)
    dev = hx.device
    eorder = None if edges_sorted else tseg.order
    he = _bytes(lib().value("wsis_ecc_he_bytes", E), dev)
    lib().call("wsis_ecc_edge_mlp", _ptr(edgefeats), _ptr(eorder), E, _ptr(mlp), _ptr(he), _stream())
    msg = torch.empty((max(E, 1), F), dtype=torch.float32, device=dev)
    width = F * (nrepeats + 1)
    cat = torch.empty((S, width), dtype=torch.float32, device=dev) if cat_all else None
    if cat_all:
        cat[:, :F] = hx
    h = hx
    for r in range(nrepeats):
        out = torch.empty_like(h)
        lib().call("wsis_ecc_messages", _ptr(he), _ptr(w4p), _ptr(b4), _ptr(h), _ptr(src), _ptr(eorder), E, _ptr(msg), _stream())
        cat_ptr = ctypes.c_void_p(cat.data_ptr() + 4 * F * (r + 1)) if cat_all else None
        lib().call("wsis_ecc_gru_step_msg", _ptr(h), _ptr(msg), _ptr(tseg.offsets), S, _ptr(params), int(layernorm),
                   float(eps), _ptr(out), cat_ptr, width, _stream())
        h = out
    return cat if cat_all else h


def pack_pos_mlp(fc_position):
    """fc_position = Sequential(Linear(3,16), ReLU, Linear(16,1)) (backbone_3D_WSIS.py:110-114) -> f32[81]."""
    l1, l2 = fc_position[0], fc_position[2]
    return torch.cat([l1.weight.detach().reshape(-1), l1.bias.detach().reshape(-1), l2.weight.detach().reshape(-1),
                      l2.bias.detach().reshape(-1)]).float().contiguous()


def random_walk(edge_u, edge_v, affinity, seed_label, pred, conf, class_num, iterations, useg=None, vseg=None):
    """weak_label_propagation (scannetv2_dataset.py:664-735) on the device, float64.
    Returns (pseudo int32[S] with -100 = none, score float64[S])."""
    edge_u = _cuda(edge_u, "edge_u").contiguous()
    edge_v = edge_v.contiguous()
    S = seed_label.shape[0]
    E = edge_u.shape[0]
    dev = edge_u.device
    useg = useg or SegmentIndex(edge_u, S)
    vseg = vseg or SegmentIndex(edge_v, S)
    pseudo = torch.empty((S,), dtype=torch.int32, device=dev)
    score = torch.empty((S,), dtype=torch.float64, device=dev)
    ws = _bytes(lib().value("wsis_random_walk_ws_bytes", S, int(class_num)), dev)
    # converted copies are bound to names so that they outlive the launch call (a temporary freed right after
    # _ptr() would be handed to the next temporary by the caching allocator)
    aff32 = affinity.contiguous().float()
    seed32, pred32, conf32 = seed_label.int().contiguous(), pred.int().contiguous(), conf.float().contiguous()
    lib().call("wsis_random_walk", _ptr(edge_u), _ptr(edge_v), _ptr(aff32), _ptr(useg.order), _ptr(useg.offsets),
               _ptr(vseg.order), _ptr(vseg.offsets), S, E, _ptr(seed32), _ptr(pred32), _ptr(conf32), int(class_num),
               int(iterations), _ptr(pseudo), _ptr(score), _ptr(ws), _stream())
    return pseudo, score
