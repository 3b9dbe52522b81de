"""TEST / BASELINE INFRASTRUCTURE ONLY -- the hot path on the host CPU, for bench.py's `cpu_baseline` and
`--impl reference` legs and for end-to-end parity checks.

Sparse rulebooks and convolutions run on the UNMODIFIED reference spconv CPU kernels (oracle/_ref, compiled from
/root/reference by oracle/Makefile; kind = "reference"), or on the plain-C restatement when that library is
absent (kind = "port").  Everything the reference obtains from libraries that are not in its tree
(pointgroup_ops, torch_scatter, torch_geometric) uses the restatements of oracle/oracle.py and torch CPU ops,
as the BASELINE.md plan prescribes.  The weights are read from a network object built by the product's host
mirror (same parameter names as the reference); no product kernel runs here.

Follows train_scannetv2.py:149-198 -> backbone_3D_WSIS.py:164-255 -> sparse_unet3d.py:163-172,321-350.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc
from . import ref_spconv


class _RefConvFn(torch.autograd.Function):
    """Autograd over the reference's CPU kernels: forward = indiceConv, backward = indiceConvBackward
    (spconv_ops.h:253-433), exactly the pairing of modules/lib/spconv/spconv/functional.py:21-117."""

    @staticmethod
    def forward(ctx, feats, weight, pairs, num, n_out, inverse, subm):
        ctx.save_for_backward(feats, weight, pairs, num)
        ctx.inverse, ctx.subm = inverse, subm
        return ref_spconv.indice_conv(feats, weight, pairs, num, n_out, inverse, subm)

    @staticmethod
    def backward(ctx, g):
        feats, weight, pairs, num = ctx.saved_tensors
        din, dw = ref_spconv.indice_conv_backward(feats, weight, g.contiguous(), pairs, num, ctx.inverse, ctx.subm)
        return din, dw, None, None, None, None, None


class CpuSpconv:
    """get_indice_pairs + indice_conv on the CPU with a per-key rulebook cache (conv.py:140-152)."""

    def __init__(self, prefer_reference=True, train=False):
        self.train = train
        self.use_ref = prefer_reference and ref_spconv.available()
        self.kind = "reference" if self.use_ref else "port"
        self.rulebooks = {}
        self.t_rulebook = 0.0
        self.t_conv = 0.0

    def rulebook(self, key, coords, batch_size, shape, ksize, stride, padding, subm):
        if key in self.rulebooks:
            return self.rulebooks[key]
        t0 = time.perf_counter()
        if self.use_ref:
            outids, pairs, num = ref_spconv.get_indice_pairs(coords, batch_size, shape, ksize, stride, padding, 1, 0, subm)
            oshape = shape if subm else [(shape[i] + 2 * padding - (ksize - 1) - 1) // stride + 1 for i in range(3)]
        else:
            c = coords.numpy()
            if subm:
                p, n = orc.rulebook_subm(c, batch_size, shape, ksize, 1)
                outids, pairs, num, oshape = coords, torch.from_numpy(p), torch.from_numpy(n), shape
            else:
                oc, p, n, oshape = orc.rulebook_conv(c, batch_size, shape, ksize, stride, padding, 1)
                outids, pairs, num = torch.from_numpy(oc), torch.from_numpy(p), torch.from_numpy(n)
        self.t_rulebook += time.perf_counter() - t0
        self.rulebooks[key] = (outids, pairs, num, list(oshape))
        return self.rulebooks[key]

    def conv(self, feats, weight, pairs, num, n_out, inverse=False, subm=False):
        t0 = time.perf_counter()
        if self.train:
            assert self.use_ref, "the CPU training step needs oracle/_ref (the reference's backward kernels)"
            out = _RefConvFn.apply(feats, weight, pairs, num, n_out, inverse, subm)
        elif self.use_ref:
            out = ref_spconv.indice_conv(feats, weight, pairs, num, n_out, inverse, subm)
        else:
            out = torch.from_numpy(orc.indice_conv(feats.numpy(), weight.numpy(), pairs.numpy(), num.numpy(), n_out, inverse))
        self.t_conv += time.perf_counter() - t0
        return out


def _bn_relu(bn, x):
    if bn.training:      # training step: the module itself (batch statistics, running-stat update), as the reference does
        return F.relu(bn(x))
    return F.relu(F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))


def _residual_block(sp, blk, x, coords, bs, shape, key):
    mods = list(blk.conv_branch._modules.values())
    bn1, conv1, bn2, conv2 = mods[0], mods[2], mods[3], mods[5]
    ib = list(blk.i_branch._modules.values())[0]
    ident = x if isinstance(ib, torch.nn.Identity) else x @ ib.weight.view(ib.in_channels, ib.out_channels)
    _, pairs, num, _ = sp.rulebook(key, coords, bs, shape, 3, 1, 1, True)
    y = sp.conv(_bn_relu(bn1, x), conv1.weight, pairs, num, x.shape[0], False, True)
    y = sp.conv(_bn_relu(bn2, y), conv2.weight, pairs, num, x.shape[0], False, True)
    return y + ident


def _ublock(sp, ub, x, coords, bs, shape, level):
    for blk in ub.blocks._modules.values():
        x = _residual_block(sp, blk, x, coords, bs, shape, "subm%d" % level)
    if len(ub.nPlanes) > 1:
        bn, _, down = list(ub.conv._modules.values())
        outids, pairs, num, oshape = sp.rulebook("spconv%d" % level, coords, bs, shape, 2, 2, 0, False)
        y = sp.conv(_bn_relu(bn, x), down.weight, pairs, num, outids.shape[0], False, False)
        y = _ublock(sp, ub.u, y, outids, bs, oshape, level + 1)
        bn, _, up = list(ub.deconv._modules.values())
        y = sp.conv(_bn_relu(bn, y), up.weight, pairs, num, x.shape[0], True, False)
        x = torch.cat((x, y), dim=1)
        for blk in ub.blocks_tail._modules.values():
            x = _residual_block(sp, blk, x, coords, bs, shape, "subm%d" % level)
    return x


def _scatter(src, index, reduce, S):
    shape = (S,) + tuple(src.shape[1:])
    if reduce == "max":
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return torch.zeros(shape, dtype=src.dtype).scatter_reduce(0, idx, src, "amax", include_self=False)
    out = torch.zeros(shape, dtype=src.dtype).index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(S, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
        out = out / cnt.view(-1, *([1] * (src.dim() - 1)))
    return out


def forward(net, batch, prefer_reference=True, keep=None):
    with torch.no_grad():
        return _forward(net, batch, prefer_reference, keep, False)


def train_step(net, batch, criterion, epoch=121):
    """One CPU training step body (train_scannetv2.py:149-244: forward, MultiTaskLoss, backward) on the reference's
    CPU kernels with torch BatchNorm in training mode.  `net` = a CPU, train-mode network; `criterion` = a
    MultiTaskLoss.  Returns (loss, loss parts, ret, stage timings).  Gradients are left in the parameters' .grad."""
    T0 = time.perf_counter()
    ret, T, kind = _forward(net, batch, True, None, True)
    t0 = time.perf_counter()
    from wsis_b200.train import loss_inputs           # pure dict plumbing (train_scannetv2.py:211-231)
    loss, parts = criterion(loss_inputs(ret, batch), epoch)
    T["loss"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    loss.backward()
    T["backward"] = time.perf_counter() - t0
    T["total"] = time.perf_counter() - T0
    return loss.detach(), parts, ret, T


def _forward(net, batch, prefer_reference, keep, train):
    """CPU forward of one collated batch with `net` (a CPU, eval-mode wsis_b200.model.Network).
    Returns (ret dict of torch CPU tensors, stage timings dict, kind).  When `keep` is a dict it receives what the
    parity checks compare besides the outputs: the nine rulebooks (key -> (in coords, out coords, pairs, num)), the
    voxelization maps and the U-Net output features."""
    sp = CpuSpconv(prefer_reference, train)
    T = {}
    t0 = time.perf_counter()
    voxel_locs, p2v, v2p = (torch.from_numpy(a) for a in orc.voxelization_idx(batch["locs"].numpy(), batch["batch_size"], 4))
    T["voxelization_idx"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    S = batch["num_superpoints"]
    superpoint = batch["superpoint"]
    centers = _scatter(batch["locs_float"], superpoint, "mean", S)
    feats = torch.cat((batch["feats"], batch["locs_float"]), 1)
    voxel_feats = torch.from_numpy(orc.voxelization(feats.numpy(), v2p.numpy(), 4))
    T["voxelization"] = time.perf_counter() - t0
    coords, bs, shape = voxel_locs.int(), batch["batch_size"], list(batch["spatial_shape"])
    t0 = time.perf_counter()
    conv0 = list(net.input_conv._modules.values())[0]
    _, pairs, num, _ = sp.rulebook("subm1", coords, bs, shape, 3, 1, 1, True)
    x = sp.conv(voxel_feats, conv0.weight, pairs, num, coords.shape[0], False, True)
    x = _ublock(sp, net.unet, x, coords, bs, shape, 1)
    x = _bn_relu(list(net.output_layer._modules.values())[0], x)
    T["unet"] = time.perf_counter() - t0
    T["unet_rulebooks"], T["unet_convs"] = sp.t_rulebook, sp.t_conv
    if keep is not None:
        keep.update(rulebooks=dict(sp.rulebooks), voxel_locs=voxel_locs, p2v=p2v, v2p=v2p, unet_out=x)
    t0 = time.perf_counter()
    output_feats = x[p2v.long()]
    ret = {"semantic_scores": net.linear(output_feats)}
    embeddings = _scatter(output_feats, superpoint, "mean", S)
    T["gather_pool_head"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    from types import SimpleNamespace  # tensor-only graph info with the interface the ECC module expects
    gi = SimpleNamespace(get_buffers=lambda: (None, None, None, None, batch["ecc_edgefeats"]),
                         get_pyg_buffers=lambda: batch["ecc_edge_index"], cuda=lambda: None)
    net.ecc.set_info([gi], cuda=False)
    ecc = net.ecc(embeddings)
    T["ecc_gru"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ret['sp_semantic_scores'] = net.sp_sem_seg(ecc)
    ret['pred_sp_offset_vectors'] = net.sp_offset_vector_head(ecc)
    ret['pred_sp_occupancy'] = net.sp_occupancy_head(ecc).squeeze(-1)
    ret['pred_sp_ins_size'] = net.sp_ins_size_head(ecc).squeeze(-1)
    q, k, v = net.w_qs(ecc), net.w_ks(ecc), net.w_vs(ecc)
    eu, ev = batch["edge_u_list"], batch["edge_v_list"]
    pos = net.fc_position(centers[eu] - centers[ev]).reshape(-1)
    a = (q[eu] * k[ev]).sum(1) / np.sqrt(k.size(-1)) * pos
    a = a - _scatter(a, eu, "max", S)[eu]
    ea = torch.exp(a)
    a = ea / _scatter(ea, eu, "sum", S)[eu]
    ret['edge_affinity'] = a
    sp_feat = ecc + _scatter(a.reshape(-1, 1) * v[ev], eu, "sum", S)
    ret['sp_discriminative_feats'] = net.feature_term(sp_feat)
    T["heads_attention"] = time.perf_counter() - t0
    return ret, T, sp.kind
