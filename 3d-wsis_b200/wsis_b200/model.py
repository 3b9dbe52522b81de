"""Host-side mirror of the reference model for the hot path (built on the drop-in `spconv` package).

The reference's model files are pure operator-API usage; on a machine that has the
reference tree they run UNMODIFIED on top of `3d-wsis_b200/spconv` + `pointgroup_ops`.  This mirror exists so
that the same network can be constructed where the reference tree is absent (the GPU box) and so that the
inference path can use the fused kernels (BatchNorm+ReLU prologue, residual epilogue, gather+pool,
edge attention).  Module / parameter names, shapes and construction ORDER follow the reference exactly, so a
reference `state_dict` loads with strict=True and fixed-seed initialisation is identical:

    ResidualBlock, UBlock   modules/model/sparse_unet3d.py:104-172, 229-350
    GraphNetwork (ECC-GRU)  modules/model/graphnet.py:21-114, spg_modules.py:128-253
    Network                 modules/model/backbone_3D_WSIS.py:25-255
"""
import ast
import functools
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

import spconv
from spconv.modules import SparseModule, _fold_bn

from . import ops as W


# ---------------------------------------------------------------------------------------------------------
# sparse U-Net (sparse_unet3d.py)
# ---------------------------------------------------------------------------------------------------------
class ResidualBlock(SparseModule):
    """Pre-norm residual block: BN-ReLU-SubM3 -BN-ReLU-SubM3 + identity (1x1 SubM when widths differ).
    sparse_unet3d.py:104-172."""

    def __init__(self, in_channels, out_channels, norm_fn, indice_key=None, normalize_before=True):
        super().__init__()
        if in_channels == out_channels:
            self.i_branch = spconv.SparseSequential(nn.Identity())
        else:
            self.i_branch = spconv.SparseSequential(
                spconv.SubMConv3d(in_channels, out_channels, kernel_size=1, bias=False))
        conv = functools.partial(spconv.SubMConv3d, kernel_size=3, padding=1, bias=False, indice_key=indice_key)
        if normalize_before:
            self.conv_branch = spconv.SparseSequential(
                norm_fn(in_channels), nn.ReLU(), conv(in_channels, out_channels),
                norm_fn(out_channels), nn.ReLU(), conv(out_channels, out_channels))
        else:
            self.conv_branch = spconv.SparseSequential(
                conv(in_channels, out_channels), norm_fn(out_channels), nn.ReLU(),
                conv(out_channels, out_channels), norm_fn(out_channels), nn.ReLU())
        self.normalize_before = normalize_before

    def _fused_ok(self, input):
        return (self.normalize_before and not torch.is_grad_enabled() and not self.training
                and input.features.is_cuda and input.features.dtype == torch.float32
                and input.indices.shape[0] != 0)

    def forward(self, input):
        identity = spconv.SparseConvTensor(input.features, input.indices, input.spatial_shape, input.batch_size)
        if self._fused_ok(input):
            # inference: 2 kernels per block (+1 small GEMM when widths differ): BN+ReLU ride in each conv's gather
            # prologue, the identity add rides in the second conv's epilogue
            bn1, _, conv1, bn2, _, conv2 = list(self.conv_branch._modules.values())
            res = self.i_branch(identity).features
            mid = conv1(input, _prologue=_fold_bn(bn1) + (1,))
            return conv2(mid, _prologue=_fold_bn(bn2) + (1,), _residual=res)
        if (self.normalize_before and self.training and torch.is_grad_enabled() and input.features.is_cuda
                and input.features.dtype == torch.float32 and input.indices.shape[0] > 1):
            # training: batch-statistics BN+ReLU in each conv's prologue, the identity add in the second conv's epilogue
            from . import train as T
            bn1, _, conv1, bn2, _, conv2 = list(self.conv_branch._modules.values())
            if T.FUSED and bn1.training and bn2.training:
                res = self.i_branch(identity).features
                mid = conv1(input, _bn_train=bn1)
                return conv2(mid, _bn_train=bn2, _residual=res)
        output = self.conv_branch(input)
        output.features += self.i_branch(identity).features
        return output


class UBlock(nn.Module):
    """Recursive encoder/decoder level.  sparse_unet3d.py:229-350."""

    def __init__(self, nPlanes, norm_fn, block_reps=2, block=ResidualBlock, indice_key_id=1, normalize_before=True):
        super().__init__()
        self.nPlanes = nPlanes
        self.blocks = spconv.SparseSequential(OrderedDict(
            (f"block{i}", block(nPlanes[0], nPlanes[0], norm_fn, normalize_before=normalize_before,
                                indice_key=f"subm{indice_key_id}")) for i in range(block_reps)))
        if len(nPlanes) > 1:
            down = spconv.SparseConv3d(nPlanes[0], nPlanes[1], kernel_size=2, stride=2, bias=False,
                                       indice_key=f"spconv{indice_key_id}")
            if normalize_before:
                self.conv = spconv.SparseSequential(norm_fn(nPlanes[0]), nn.ReLU(), down)
            else:
                self.conv = spconv.SparseSequential(down, norm_fn(nPlanes[1]), nn.ReLU())
            self.u = UBlock(nPlanes[1:], norm_fn, block_reps, block, indice_key_id=indice_key_id + 1,
                            normalize_before=normalize_before)
            up = spconv.SparseInverseConv3d(nPlanes[1], nPlanes[0], kernel_size=2, bias=False,
                                            indice_key=f"spconv{indice_key_id}")
            if normalize_before:
                self.deconv = spconv.SparseSequential(norm_fn(nPlanes[1]), nn.ReLU(), up)
            else:
                self.deconv = spconv.SparseSequential(up, norm_fn(nPlanes[0]), nn.ReLU())
            self.blocks_tail = spconv.SparseSequential(OrderedDict(
                (f"block{i}", block(nPlanes[0] * (2 - i), nPlanes[0], norm_fn, indice_key=f"subm{indice_key_id}",
                                    normalize_before=normalize_before)) for i in range(block_reps)))

    def forward(self, input):
        output = self.blocks(input)
        identity = spconv.SparseConvTensor(output.features, output.indices, output.spatial_shape, output.batch_size)
        if len(self.nPlanes) > 1:
            decoder = self.deconv(self.u(self.conv(output)))
            output.features = torch.cat((identity.features, decoder.features), dim=1)
            output = self.blocks_tail(output)
        return output


# ---------------------------------------------------------------------------------------------------------
# ECC-GRU graph network on the superpoint graph (graphnet.py, spg_modules.py) -- plain torch ops for now
# (a fused graph kernel is the next row after the UNet path)
# ---------------------------------------------------------------------------------------------------------
class GraphInfo(object):
    """Tensor-only stand-in for ecc.GraphConvInfo (ecc/GraphConvInfo.py:16-104): `edge_index` int64[2,E] =
    (source, target) sorted by target, `edgefeats` f32[E,13] in the same order."""

    def __init__(self, edge_index, edgefeats):
        self._edge_indexes = edge_index
        self._edgefeats = edgefeats

    def cuda(self):
        self._edge_indexes = self._edge_indexes.cuda()
        self._edgefeats = self._edgefeats.cuda()
        return self

    def to(self, device):
        self._edge_indexes = self._edge_indexes.to(device)
        self._edgefeats = self._edgefeats.to(device)
        return self

    def get_buffers(self):
        return None, None, None, None, self._edgefeats

    def get_pyg_buffers(self):
        return self._edge_indexes


def create_fnet(widths, orthoinit, llbias, bnidx=-1):
    """Filter-generating MLP (graphnet.py:21-39)."""
    mods = []
    for k in range(len(widths) - 2):
        mods.append(nn.Linear(widths[k], widths[k + 1]))
        if orthoinit:
            init.orthogonal_(mods[-1].weight, gain=init.calculate_gain('relu'))
        if bnidx == k:
            mods.append(nn.BatchNorm1d(widths[k + 1]))
        mods.append(nn.ReLU(True))
    mods.append(nn.Linear(widths[-2], widths[-1], bias=llbias))
    if orthoinit:
        init.orthogonal_(mods[-1].weight)
    if bnidx == len(widths) - 1:
        mods.append(nn.BatchNorm1d(mods[-1].weight.size(0)))
    return nn.Sequential(*mods)


class GRUCellEx(nn.GRUCell):
    """GRU cell with un-affine layer normalisation of both gate pre-activations and an input gate
    (spg_modules.py:208-253)."""

    def __init__(self, input_size, hidden_size, bias=True, layernorm=True, ingate=True):
        super().__init__(input_size, hidden_size, bias)
        self._layernorm = layernorm
        self._ingate = ingate
        if layernorm:
            self.add_module('ini', nn.InstanceNorm1d(1, eps=1e-5, affine=False, track_running_stats=False))
            self.add_module('inh', nn.InstanceNorm1d(1, eps=1e-5, affine=False, track_running_stats=False))
        if ingate:
            self.add_module('ig', nn.Linear(hidden_size, input_size, bias=True))

    def forward(self, input, hidden):
        if self._ingate:
            input = torch.sigmoid(self._modules['ig'](hidden)) * input
        gi = F.linear(input, self.weight_ih)
        gh = F.linear(hidden, self.weight_hh)
        if self._layernorm:
            # InstanceNorm1d(1, affine=False, no running stats) over [S, 1, L] normalises every row over L with the
            # biased variance and eps = 1e-5, which is exactly an un-affine layer norm over the last dimension; the
            # instance-norm kernel costs ~60 us per call on these tiny [S, 96] matrices, layer_norm a few us
            gi = F.layer_norm(gi, (gi.shape[-1],), None, None, self._modules['ini'].eps)
            gh = F.layer_norm(gh, (gh.shape[-1],), None, None, self._modules['inh'].eps)
        i_r, i_i, i_n = gi.chunk(3, 1)
        h_r, h_i, h_n = gh.chunk(3, 1)
        bih_r, bih_i, bih_n = self.bias_ih.chunk(3)
        bhh_r, bhh_i, bhh_n = self.bias_hh.chunk(3)
        resetgate = torch.sigmoid(i_r + bih_r + h_r + bhh_r)
        inputgate = torch.sigmoid(i_i + bih_i + h_i + bhh_i)
        newgate = torch.tanh(i_n + bih_n + resetgate * (h_n + bhh_n))
        return newgate + inputgate * (hidden - newgate)


class NNConv(nn.Module):
    """Parameter-free edge-conditioned convolution as the reference actually runs it (spg_modules.py:24-121 with
    aggr='mean', root_weight=False, bias=False, vv=False).  NNConv.__init__ accepts flow="target_to_source" but
    never forwards it to MessagePassing.__init__ (spg_modules.py:61-72), so PyG's default source_to_target
    applies: for edge_index = (source, target),
        out[t] = mean over edges (s,t) of  x[s]^T . W_e."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels

    def forward(self, x, edge_index, weights):
        src, tgt = edge_index[0], edge_index[1]
        msg = torch.bmm(x[src].unsqueeze(1), weights.view(-1, self.in_channels, self.out_channels)).squeeze(1)
        out = torch.zeros((x.shape[0], self.out_channels), dtype=x.dtype, device=x.device).index_add_(0, tgt, msg)
        cnt = torch.zeros((x.shape[0],), dtype=x.dtype, device=x.device).index_add_(
            0, tgt, torch.ones_like(tgt, dtype=x.dtype))
        return out / cnt.clamp(min=1).unsqueeze(1)


class RNNGraphConvModule(nn.Module):
    """spg_modules.py:128-185 with use_pyg=True."""

    def __init__(self, cell, filter_net, nfeat, nrepeats=1, cat_all=False):
        super().__init__()
        self._cell = cell
        self._fnet = filter_net
        self._nrepeats = nrepeats
        self._cat_all = cat_all
        self._gci = None
        self.nn = NNConv(nfeat, nfeat)

    def set_info(self, gc_info):
        self._gci = gc_info

    def forward(self, hx):
        edgefeats = self._gci.get_buffers()[4]
        edge_index = self._gci.get_pyg_buffers()
        nc = hx.size(1)
        cell = self._cell
        fast = (not torch.is_grad_enabled() and hx.is_cuda and hx.dtype == torch.float32 and nc == 32
                and isinstance(cell, GRUCellEx) and cell._ingate and cell.bias)
        if fast and W.ECC_FUSED_FILTERS and W.ecc_fnet_supported(self._fnet) and edgefeats.dtype == torch.float32:
            # inference: the [E,1024] filters are never materialised (csrc/ecc_umma.cu regenerates them on the tensor
            # cores inside every step), one message kernel + one GRU kernel per step
            key = tuple(p._version for p in cell.parameters()) + tuple(p.data_ptr() for p in cell.parameters()) + \
                (W.cache_epoch(),)
            if getattr(self, "_packed_key", None) != key:
                self._packed_key, self._packed = key, W.pack_ecc_gru(cell)
            tseg = W.SegmentIndex(edge_index[1], hx.shape[0])
            return W.ecc_gru_fused(hx, self._fnet, edgefeats, edge_index[0], tseg, self._packed, self._nrepeats,
                                   layernorm=cell._layernorm, eps=cell._modules['ini'].eps if cell._layernorm else 1e-5,
                                   cat_all=self._cat_all)
        weights = _seq(self._fnet, edgefeats)
        assert hx.dim() == 2 and weights.dim() == 2 and weights.size(1) == nc * nc
        if fast:
            # inference: one kernel per GRU step (message, gates, layer norms and the concatenation fused;
            # csrc/ecc.cu) instead of ~20 torch launches
            key = tuple(p._version for p in cell.parameters()) + tuple(p.data_ptr() for p in cell.parameters()) + \
                (W.cache_epoch(),)
            if getattr(self, "_packed_key", None) != key:
                self._packed_key, self._packed = key, W.pack_ecc_gru(cell)
            tseg = W.SegmentIndex(edge_index[1], hx.shape[0])
            return W.ecc_gru(hx, weights, edge_index[0], tseg, self._packed, self._nrepeats,
                             layernorm=cell._layernorm, eps=cell._modules['ini'].eps if cell._layernorm else 1e-5,
                             cat_all=self._cat_all)
        hxs = [hx]
        for _ in range(self._nrepeats):
            hx = self._cell(self.nn(hx, edge_index, weights), hx)
            hxs.append(hx)
        return torch.cat(hxs, 1) if self._cat_all else hx


class GraphNetwork(nn.Module):
    """graphnet.py:42-114 for the layer tokens 3D-WSIS uses ('gru_R_0', 'f_W', 'b', 'r', 'd_P')."""

    def __init__(self, config, nfeat, fnet_widths, fnet_orthoinit=True, fnet_llbias=True, fnet_bnidx=-1, **_unused):
        super().__init__()
        self.gconvs = []
        for d, conf in enumerate(config.split(',')):
            conf = conf.strip().split('_')
            if conf[0] == 'f':
                self.add_module(str(d), nn.Linear(nfeat, int(conf[1])))
                nfeat = int(conf[1])
            elif conf[0] == 'b':
                self.add_module(str(d), nn.BatchNorm1d(nfeat, eps=1e-5, affine=len(conf) == 1))
            elif conf[0] == 'r':
                self.add_module(str(d), nn.ReLU(True))
            elif conf[0] == 'd':
                self.add_module(str(d), nn.Dropout(p=float(conf[1]), inplace=False))
            elif conf[0] == 'gru':
                nrepeats = int(conf[1])
                vv = bool(int(conf[2])) if len(conf) > 2 else True
                layernorm = bool(int(conf[3])) if len(conf) > 3 else True
                ingate = bool(int(conf[4])) if len(conf) > 4 else True
                cat_all = bool(int(conf[5])) if len(conf) > 5 else True
                if vv:
                    raise NotImplementedError("vector-valued ECC filters are not used by 3D-WSIS ('gru_7_0')")
                fnet = create_fnet(fnet_widths + [nfeat ** 2], fnet_orthoinit, fnet_llbias, fnet_bnidx)
                cell = GRUCellEx(nfeat, nfeat, bias=True, layernorm=layernorm, ingate=ingate)
                gconv = RNNGraphConvModule(cell, fnet, nfeat, nrepeats=nrepeats, cat_all=cat_all)
                self.add_module(str(d), gconv)
                self.gconvs.append(gconv)
                if cat_all:
                    nfeat *= nrepeats + 1
            elif len(conf[0]) > 0:
                raise NotImplementedError('Unknown module: ' + conf[0])

    def set_info(self, gc_infos, cuda):
        gc_infos = gc_infos if isinstance(gc_infos, (list, tuple)) else [gc_infos]
        for i, gc in enumerate(self.gconvs):
            if cuda:
                gc_infos[i].cuda()
            gc.set_info(gc_infos[i])

    def forward(self, input):
        return _seq(self, input)


# ---------------------------------------------------------------------------------------------------------
# Network (backbone_3D_WSIS.py)
# ---------------------------------------------------------------------------------------------------------
def _seq(seq, x):
    """nn.Sequential forward; under autograd on CUDA, training-mode BatchNorm1d (+ReLU) runs on the batch-statistics
    kernels of wsis_b200.train (synchronised across ranks like SyncBatchNorm); in inference a Linear-BN-ReLU-Linear head
    is one kernel (csrc/heads.cu)."""
    if not torch.is_grad_enabled() and W.mlp_head_supported(seq, x):
        return W.mlp_head(seq, x)
    if torch.is_grad_enabled() and x.is_cuda:
        from . import train as T
        return T.run_sequential(seq, x)
    for module in seq._modules.values():
        x = module(x)
    return x


def _packed_of(module):
    """Per-module cache of the pre-swizzled weight image (ops.PackedWeights)."""
    pk = getattr(module, "_wsis_packed", None)
    if pk is None:
        pk = W.PackedWeights()
        module._wsis_packed = pk
    return pk


def _head(norm_fn, width, out):
    return nn.Sequential(nn.Linear(width, width, bias=True), norm_fn(width), nn.ReLU(inplace=True),
                         nn.Linear(width, out))


class Network(nn.Module):
    """backbone_3D_WSIS.py:25-255.  `param` needs input_channel, use_coords, blocks, block_reps, media, classes."""

    def __init__(self, param):
        super().__init__()
        self.input_channel = param.input_channel
        self.use_coords = param.use_coords
        self.blocks = param.blocks
        self.block_reps = param.block_reps
        self.media = param.media
        self.classes = param.classes
        if self.use_coords:
            self.input_channel += 3
        self.input_conv = spconv.SparseSequential(
            spconv.SubMConv3d(self.input_channel, self.media, kernel_size=3, padding=1, bias=False, indice_key="subm1"))
        norm_fn = functools.partial(nn.BatchNorm1d, eps=1e-4, momentum=0.1)
        self.unet = UBlock([self.media * (i + 1) for i in range(self.blocks)], norm_fn, self.block_reps, ResidualBlock,
                           indice_key_id=1)
        self.output_layer = spconv.SparseSequential(norm_fn(self.media), nn.ReLU(inplace=True))
        self.linear = _head(norm_fn, self.media, self.classes)
        self.ecc = GraphNetwork('gru_7_0,f_64,b,r', nfeat=self.media, fnet_widths=[13] + [32, 128, 64],
                                fnet_orthoinit=True, fnet_llbias=True, fnet_bnidx=2)
        d = 64
        self.sp_sem_seg = _head(norm_fn, d, self.classes)
        self.sp_offset_vector_head = _head(norm_fn, d, 3)
        self.sp_occupancy_head = _head(norm_fn, d, 1)
        self.sp_ins_size_head = _head(norm_fn, d, 1)
        self.fc_position = nn.Sequential(nn.Linear(3, 16), nn.ReLU(), nn.Linear(16, 1))
        self.w_qs = nn.Linear(d, d, bias=False)
        self.w_ks = nn.Linear(d, d, bias=False)
        self.w_vs = nn.Linear(d, d, bias=False)
        self.feature_term = _head(norm_fn, d, 7)
        # frozen sub-modules (backbone_3D_WSIS.py:35,131-136): eval() + requires_grad=False at construction and eval()
        # again on every forward (:172-173)
        fm = getattr(param, "fix_module", "[]")
        self.fix_module = ast.literal_eval(fm) if isinstance(fm, str) else list(fm or [])
        for name in self.fix_module:
            module = getattr(self, name)
            module.eval()
            for prm in module.parameters():
                prm.requires_grad = False

    def load_state_dict(self, *args, **kwargs):
        W.invalidate_caches()   # derived parameter images (packed weights, folded BN) are rebuilt at the next forward
        return super().load_state_dict(*args, **kwargs)

    def forward(self, input, input_map, extra_data):
        """Same inputs and the same result dict as backbone_3D_WSIS.py:164-255.  Optional extra_data keys that
        avoid host syncs / rebuilds: "sp_index" (ops.SegmentIndex of `superpoint`), "edge_index_u" (SegmentIndex of
        edge_u_list), "num_superpoints"."""
        ret = {}
        for name in self.fix_module:
            getattr(self, name).eval()
        output = self.output_layer(self.unet(self.input_conv(input)))
        if extra_data.get("keep_unet_features"):      # parity checks compare the U-Net output itself
            extra_data["unet_features"] = output.features
        fused = not torch.is_grad_enabled() and not self.training and output.features.is_cuda

        superpoint = extra_data["superpoint"].long()
        output_feats = None
        if fused:
            p2v = input_map if input_map.dtype == torch.int32 else input_map.int()
        elif torch.is_grad_enabled() and output.features.is_cuda and output.features.dtype == torch.float32:
            from . import train as T                                                  # training: own gather kernel + CSR backward
            p2v = input_map if input_map.dtype == torch.int32 else input_map.int()
            output_feats = T.gather_rows(output.features, p2v)                        # :179 voxel -> point
        else:
            output_feats = output.features[input_map.long()]                          # :179 voxel -> point
        if fused and W.mlp_head_supported(self.linear, output.features):
            # :179 + :182 in one kernel; the [N_points, 32] gathered features are never materialised (the pooling
            # below gathers through p2v as well)
            ret["semantic_scores"] = W.mlp_head(self.linear, output.features, gather=p2v)
        else:
            if output_feats is None:
                output_feats = W.gather_rows(output.features, p2v)
            ret["semantic_scores"] = _seq(self.linear, output_feats)                  # :182

        if fused:
            seg = extra_data.get("sp_index")
            if seg is None:
                S = extra_data.get("num_superpoints")
                S = int(superpoint.max().item()) + 1 if S is None else int(S)
                seg = W.SegmentIndex(superpoint, S)
            if output_feats is None:                                                  # :179 + :188 superpoint pooling
                embeddings = W.segment_reduce(output.features, seg, "mean", gather=p2v)
            else:
                embeddings = W.segment_reduce(output_feats, seg, "mean")
        elif torch.is_grad_enabled() and output_feats.is_cuda and output_feats.dtype == torch.float32 \
                and (extra_data.get("sp_index") is not None or extra_data.get("num_superpoints") is not None):
            from . import train as T
            seg = extra_data.get("sp_index") or W.SegmentIndex(superpoint, int(extra_data["num_superpoints"]))
            embeddings = T.segment_mean(output_feats, seg)                            # :188, no host sync, no atomics
        else:
            embeddings = _scatter_torch(output_feats, superpoint, "mean")

        self.ecc.set_info(extra_data['GIs'], cuda=output.features.is_cuda)
        ecc_outputs = self.ecc(embeddings)                                            # :191-193

        ret['sp_semantic_scores'] = _seq(self.sp_sem_seg, ecc_outputs)
        ret['pred_sp_offset_vectors'] = _seq(self.sp_offset_vector_head, ecc_outputs)
        ret['pred_sp_occupancy'] = _seq(self.sp_occupancy_head, ecc_outputs).squeeze(-1)
        ret['pred_sp_ins_size'] = _seq(self.sp_ins_size_head, ecc_outputs).squeeze(-1)

        centers = extra_data['superpoint_cenetr_xyz']
        q = k = v = None
        if fused:                                                                     # rows @ W^T on the tensor-core kernel
            q, k, v = (W.dense_rows(ecc_outputs, lin.weight, packed=_packed_of(lin), holder=ecc_outputs, linear_layout=True)
                       for lin in (self.w_qs, self.w_ks, self.w_vs))
        if q is None or k is None or v is None:
            q, k, v = self.w_qs(ecc_outputs), self.w_ks(ecc_outputs), self.w_vs(ecc_outputs)
        edge_u, edge_v = extra_data["edge_u_list"], extra_data["edge_v_list"]
        if fused:
            eseg = extra_data.get("edge_index_u") or W.SegmentIndex(edge_u, ecc_outputs.shape[0])
            affinity, sp_feat = W.edge_attention(q, k, v, ecc_outputs, centers, edge_u, edge_v, eseg,
                                                 W.pack_pos_mlp(self.fc_position))   # :209-249 in one kernel
        else:
            affinity, sp_feat = _edge_attention_torch(q, k, v, ecc_outputs, centers, edge_u, edge_v, self.fc_position)
        ret['edge_affinity'] = affinity
        ret['sp_discriminative_feats'] = _seq(self.feature_term, sp_feat)
        return ret


def _scatter_torch(src, index, reduce):
    """Autograd-capable torch formulation of torch_scatter.scatter(dim=0) for the training path."""
    S = int(index.max().item()) + 1
    shape = (S,) + tuple(src.shape[1:])
    if reduce == "max":
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_reduce(0, idx, src, "amax",
                                                                                     include_self=False)
    out = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros((S,), dtype=src.dtype, device=src.device).index_add_(
            0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
        out = out / cnt.view(-1, *([1] * (src.dim() - 1)))
    return out


def _edge_attention_torch(q, k, v, ecc_outputs, centers, edge_u, edge_v, fc_position):
    """Training-path (autograd) formulation of backbone_3D_WSIS.py:209-249."""
    pos_enc = fc_position(centers[edge_u] - centers[edge_v]).reshape(-1)
    affinity = (q[edge_u] * k[edge_v]).sum(dim=1) / np.sqrt(k.size(-1)) * pos_enc
    affinity = affinity - _scatter_torch(affinity, edge_u, "max")[edge_u]
    exp_affinity = torch.exp(affinity)
    affinity = exp_affinity / _scatter_torch(exp_affinity, edge_u, "sum")[edge_u]
    res = _scatter_torch(affinity.reshape(-1, 1) * v[edge_v], edge_u, "sum")
    sp_feat = torch.zeros(ecc_outputs.shape).to(ecc_outputs) + ecc_outputs
    sp_feat[:res.shape[0]] += res
    return affinity, sp_feat
