"""Training step of the hot path (BASELINE.json configs[2]): forward + backward of the network on the CUDA kernels,
the multi-task loss, one gradient all-reduce over NCCL and a fused AdamW update.

    reference: train_scannetv2.py:149-252 (step body), :734-738 (SyncBatchNorm + DistributedDataParallel),
               modules/model/losses_3D_WSIS.py:43-151 (MultiTaskLoss), :157-230 (discriminative loss)

What is different from running the reference's modules under autograd:
  * BatchNorm (batch statistics) + ReLU in front of a sparse conv is never materialised: `wsis_bn_stats` reduces the
    columns, the conv applies scale/shift/ReLU while it gathers rows (forward AND weight gradient), and the backward
    is the conv's dgrad followed by one reduction and one elementwise pass (`wsis_bn_bwd_reduce/apply`).  With more
    than one rank the statistics are synchronised like torch.nn.SyncBatchNorm (train_scannetv2.py:736): one tiny
    all-reduce of the fp64 column sums per BatchNorm and direction.
  * gradients accumulate into ONE flat fp32 bucket (dist.GradBucket) that is all-reduced once per step; the
    1/world average, the ECC gradient clamp (train_scannetv2.py:246-249) and AdamW run as one kernel over the flat
    parameter buffer (`wsis_adamw_step`).
"""
import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops as W
from ._lib import lib
from .ops import _bytes, _ptr, _stream

# SyncBatchNorm semantics when a process group with more than one rank exists (the reference converts every
# BatchNorm, train_scannetv2.py:736).  Set False for per-rank statistics.
SYNC_BN = True
# False = the reference formulation under autograd (torch BatchNorm, unfused convs): the cross-check of the fused path
FUSED = True


# "peer": the statistics all-reduce happens INSIDE the combine kernel through NVLink peer stores into torch symmetric
# memory (csrc/train.cu bn_sync_kernel); "nccl": dist.all_reduce of the fp64 sums between two kernels.  "peer" falls back
# to "nccl" when symmetric memory cannot be set up (no P2P access, gloo group).
SYNC_BN_TRANSPORT = os.environ.get("WSIS_SYNC_BN_TRANSPORT", "peer")


def _sync():
    return SYNC_BN and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _PeerSync(object):
    """Symmetric buffer + sequence counter for the in-kernel statistics all-reduce (one per process)."""
    SLOT = 2 * 1024 + 8            # doubles per contribution: up to C = 1024 channels

    def __init__(self):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        n = 2 * self.world * self.SLOT + 2 * self.world          # data + flags (uint32 pairs packed into doubles' space)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(n, dtype=torch.float64, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.buf.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.seq = 0

    def next_seq(self):
        self.seq += 1
        if self.seq >= (1 << 31):
            raise RuntimeError("wsis_b200: statistics sequence counter overflow")
        return self.seq


_PEER = None
_PEER_FAILED = False


def _peer():
    """The process's _PeerSync, or None when the peer transport is off / unavailable."""
    global _PEER, _PEER_FAILED
    if SYNC_BN_TRANSPORT != "peer" or _PEER_FAILED or not _sync() or dist.get_backend() != "nccl":
        return None
    if _PEER is None:
        try:
            _PEER = _PeerSync()
        except Exception as e:  # noqa: BLE001  (no symmetric memory on this build / no P2P): use NCCL
            import warnings
            warnings.warn("wsis_b200: symmetric-memory statistics all-reduce unavailable (%s); using NCCL" % (e,))
            _PEER_FAILED = True
            return None
    return _PEER


def bn_forward_stats(x, bn):
    """Batch statistics of x f32[N,C] for the BatchNorm1d module `bn` (training mode): returns (stat f32[4,C] =
    mean | invstd | scale | shift, sums f64[2C+1]) and updates the running statistics like torch does."""
    N, C = x.shape
    dev = x.device
    sums = torch.empty((2 * C + 1,), dtype=torch.float64, device=dev)
    ws = _bytes(lib().value("wsis_bn_ws_bytes", N, C), dev)
    stat = torch.empty((4, C), dtype=torch.float32, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = 0.0 if bn.momentum is None else float(bn.momentum)
    if track and bn.momentum is None:      # cumulative moving average
        momentum = 1.0 / float(bn.num_batches_tracked.item() + 1)
    peer = _peer()
    if (peer is not None or not _sync()) and C <= 512:
        # one launch pair: column sums, then ONE block that combines, all-reduces over NVLink peer memory (world > 1)
        # and finalizes
        lib().call("wsis_bn_forward_sync", _ptr(x), N, C, _ptr(ws), _ptr(sums),
                   _ptr(bn.weight.detach() if bn.weight is not None else None),
                   _ptr(bn.bias.detach() if bn.bias is not None else None), float(bn.eps), momentum,
                   _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None, _ptr(stat),
                   _ptr(peer.peers) if peer else None, peer.world if peer else 1, peer.rank if peer else 0,
                   peer.next_seq() if peer else 1, peer.SLOT if peer else 0, _stream())
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return stat, sums
    lib().call("wsis_bn_stats", _ptr(x), N, C, _ptr(ws), _ptr(sums), _stream())
    if _sync():
        dist.all_reduce(sums)
    lib().call("wsis_bn_finalize", _ptr(sums), C, _ptr(bn.weight.detach() if bn.weight is not None else None),
               _ptr(bn.bias.detach() if bn.bias is not None else None), float(bn.eps), momentum,
               _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None, _ptr(stat), _stream())
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return stat, sums


def bn_backward(x, da, stat, sums_fwd, relu, want_affine=True):
    """-> (dx, dgamma, dbeta) for y = relu?(x*scale+shift) given da = dL/dy."""
    N, C = x.shape
    dev = x.device
    da = da.contiguous()
    sums = torch.empty((2 * C,), dtype=torch.float64, device=dev)
    dgamma = torch.empty((C,), dtype=torch.float32, device=dev) if want_affine else None
    dbeta = torch.empty((C,), dtype=torch.float32, device=dev) if want_affine else None
    ws = _bytes(lib().value("wsis_bn_ws_bytes", N, C), dev)
    peer = _peer() if C <= 512 else None
    if peer is not None:
        lib().call("wsis_bn_bwd_reduce_sync", _ptr(x), _ptr(da), N, C, _ptr(stat), int(relu), _ptr(ws), _ptr(sums),
                   _ptr(dgamma), _ptr(dbeta), _ptr(peer.peers), peer.world, peer.rank, peer.next_seq(), peer.SLOT, _stream())
    else:
        lib().call("wsis_bn_bwd_reduce", _ptr(x), _ptr(da), N, C, _ptr(stat), int(relu), _ptr(ws), _ptr(sums), _ptr(dgamma),
                   _ptr(dbeta), _stream())
        if _sync():
            dist.all_reduce(sums)
    dx = torch.empty_like(x)
    count = sums_fwd[2 * C:]
    lib().call("wsis_bn_bwd_apply", _ptr(x), _ptr(da), N, C, _ptr(stat), int(relu), _ptr(sums), _ptr(count), None,
               _ptr(dx), _stream())
    return dx, dgamma, dbeta


class _BNReLU(torch.autograd.Function):
    """y = relu?(batch_norm_train(x)) with the reductions on the repo's kernels (heads, output layer, ECC)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu):
        x = x.contiguous()
        stat, sums = bn_forward_stats(x, bn)
        ctx.save_for_backward(x, stat, sums)
        ctx.relu, ctx.affine = relu, gamma is not None
        return W.affine_relu(x, stat[2], stat[3], relu)

    @staticmethod
    def backward(ctx, g):
        x, stat, sums = ctx.saved_tensors
        dx, dgamma, dbeta = bn_backward(x, g, stat, sums, ctx.relu, ctx.affine)
        return dx, dgamma, dbeta, None, None


def batch_norm_train(x, bn, relu):
    return _BNReLU.apply(x, bn.weight, bn.bias, bn, bool(relu))


class _GatherRows(torch.autograd.Function):
    """out[i] = src[idx[i]] (voxel -> point gather, backbone_3D_WSIS.py:179) with the backward as a segmented sum over
    the CSR of idx (one warp per source row, fixed order, no atomics) instead of torch's index_put accumulate."""

    @staticmethod
    def forward(ctx, src, idx32):
        ctx.save_for_backward(idx32)
        ctx.n_src = src.shape[0]
        return W.gather_rows(src, idx32)

    @staticmethod
    def backward(ctx, g):
        (idx32,) = ctx.saved_tensors
        seg = getattr(idx32, "_wsis_seg", None)
        if seg is None or seg.S != ctx.n_src:
            seg = W.SegmentIndex(idx32.long(), ctx.n_src)
            try:
                idx32._wsis_seg = seg
            except Exception:  # noqa: BLE001
                pass
        return W.segment_reduce(g.contiguous(), seg, "sum"), None


class _SegmentMean(torch.autograd.Function):
    """out[s] = mean of the rows of segment s (superpoint pooling, backbone_3D_WSIS.py:188); backward
    dsrc[i] = dout[ids[i]] / count[ids[i]] as one row gather."""

    @staticmethod
    def forward(ctx, src, seg):
        ctx.seg = seg
        return W.segment_reduce(src.contiguous(), seg, "mean")

    @staticmethod
    def backward(ctx, g):
        seg = ctx.seg
        cnt = (seg.offsets[1:] - seg.offsets[:-1]).clamp(min=1).to(g.dtype).unsqueeze(1)
        ids32 = getattr(seg, "_ids32", None)
        if ids32 is None:
            ids32 = seg.ids.int()
            seg._ids32 = ids32
        return W.gather_rows((g / cnt).contiguous(), ids32), None


def gather_rows(src, idx32):
    return _GatherRows.apply(src, idx32)


def segment_mean(src, seg):
    return _SegmentMean.apply(src, seg)


class _BNReLUConv(torch.autograd.Function):
    """y = sparse_conv(relu(batch_norm_train(x))) (+ residual).  mode "fwd": submanifold / strided conv (destination =
    the rulebook's output side); mode "inv": inverse conv (destination = the couple conv's input side)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, weight, residual, bn, rb, mode, packed):
        x = x.contiguous()
        stat, sums = bn_forward_stats(x, bn)
        w3 = weight.detach().reshape(-1, weight.shape[-2], weight.shape[-1])
        K, Cin, Cout = w3.shape
        use_tiles = W.get_precision() != "simt" and K <= 32 and W.umma_supported(Cin, Cout)
        if mode == "inv":
            map_, flip, n_dst, tiles = rb.nbr_in, 0, rb.n_in, rb.tiles_in
        else:
            (map_, flip), n_dst, tiles = rb.fwd_map(), rb.n_out, rb.tiles_out
        y = W.sparse_conv(x, w3, map_, n_dst, flip, False, (stat[2], stat[3], 1), residual, packed,
                          tiles=tiles() if use_tiles else None)
        ctx.save_for_backward(x, weight, stat, sums)
        ctx.rb, ctx.mode, ctx.has_res, ctx.affine = rb, mode, residual is not None, gamma is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, stat, sums = ctx.saved_tensors
        rb, mode = ctx.rb, ctx.mode
        g = g.contiguous()
        w3 = weight.detach().reshape(-1, weight.shape[-2], weight.shape[-1])
        K, Cin, Cout = w3.shape
        use_tiles = W.get_precision() != "simt" and K <= 32 and W.umma_supported(Cout, Cin)
        if mode == "inv":
            da = W.sparse_conv(g, w3, rb.nbr_out, rb.n_out, 0, True, tiles=rb.tiles_out() if use_tiles else None)
            dw = W.sparse_conv_wgrad(x, rb.nbr_in, rb.n_in, 0, g, K, Cin, Cout, prologue=(stat[2], stat[3], 1),
                                     order=rb.order_hint("in"))
        else:
            fmap, fflip = rb.fwd_map()
            da = W.sparse_conv(g, w3, rb.nbr_in, rb.n_in, 0, True, tiles=rb.tiles_in() if use_tiles else None)
            dw = W.sparse_conv_wgrad(x, fmap, rb.n_out, fflip, g, K, Cin, Cout, prologue=(stat[2], stat[3], 1),
                                     order=rb.order_hint("out"))
        dx, dgamma, dbeta = bn_backward(x, da, stat, sums, 1, ctx.affine)
        return dx, dgamma, dbeta, dw.view_as(weight), (g if ctx.has_res else None), None, None, None, None


def bn_relu_conv(x, bn, weight, rb, mode, residual=None, packed=None):
    return _BNReLUConv.apply(x, bn.weight, bn.bias, weight, residual, bn, rb, mode, packed)


def trainable_bn(bn, x):
    """True when `bn` should run on the batch-statistics kernels: a training-mode BatchNorm1d on CUDA fp32 rows."""
    return (FUSED and isinstance(bn, nn.BatchNorm1d) and bn.training and torch.is_tensor(x) and x.is_cuda
            and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 1)


def run_sequential(seq, x):
    """nn.Sequential forward with training-mode BatchNorm1d (+ the ReLU behind it) on the repo's kernels."""
    mods = list(seq._modules.values()) if isinstance(seq, nn.Module) else list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if trainable_bn(m, x):
            relu = i + 1 < len(mods) and type(mods[i + 1]) is nn.ReLU
            x = batch_norm_train(x, m, relu)
            i += 2 if relu else 1
            continue
        x = m(x)
        i += 1
    return x


class _CEDice(torch.autograd.Function):
    """CrossEntropyLoss(ignore_index) [+ multi-class dice on the softmax of the labelled rows] (losses_3D_WSIS.py:
    52-64, 72-74) as two passes over the scores (csrc/train.cu ce_dice_*)."""

    @staticmethod
    def forward(ctx, scores, labels, ignore_label, dice):
        scores = scores.contiguous()
        labels = labels.contiguous()
        N, C = scores.shape
        cp = (C + 3) // 4 * 4
        fin = torch.empty((2 + 2 * cp,), dtype=torch.float32, device=scores.device)
        ws = _bytes(lib().value("wsis_ce_dice_ws_bytes", N, C), scores.device)
        lib().call("wsis_ce_dice_fwd", _ptr(scores), _ptr(labels), N, C, int(ignore_label), int(dice), _ptr(ws), _ptr(fin),
                   _stream())
        ctx.save_for_backward(scores, labels, fin)
        ctx.ignore = int(ignore_label)
        return fin[0].clone()

    @staticmethod
    def backward(ctx, g):
        scores, labels, fin = ctx.saved_tensors
        N, C = scores.shape
        d = torch.empty_like(scores)
        g = g.contiguous().float()
        lib().call("wsis_ce_dice_bwd", _ptr(scores), _ptr(labels), N, C, ctx.ignore, _ptr(fin), _ptr(g), _ptr(d), _stream())
        return d, None, None, None


def ce_dice_loss(scores, labels, ignore_label=-100, dice=True):
    """The fused kernel for float32 CUDA scores with at most 32 classes; the torch formulation otherwise."""
    if scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2 and scores.shape[1] <= 32 \
            and labels.dtype == torch.int64 and FUSED:
        return _CEDice.apply(scores, labels, ignore_label, bool(dice))
    loss = F.cross_entropy(scores, labels, ignore_index=ignore_label)
    if dice:
        keep = labels != ignore_label
        C = scores.shape[1]
        p = F.softmax(scores, dim=-1) * keep.unsqueeze(1)
        onehot = F.one_hot(labels.clamp(min=0), C) * keep.unsqueeze(1)
        d = (2 * (p * onehot).sum(0) + 1e-5) / ((p * p).sum(0) + onehot.sum(0) + 1e-4 + 1e-5)
        loss = loss + (1.0 - d).mean()
    return loss


# ---------------------------------------------------------------------------------------------------------
# loss (modules/model/losses_3D_WSIS.py) -- one vectorised formulation for the whole batch, no per-scene Python loop
# ---------------------------------------------------------------------------------------------------------
class MultiTaskLoss(nn.Module):
    """losses_3D_WSIS.py:13-151 with the same terms, weights and reductions.  The discriminative loss
    (:157-230) of every scene of the batch is evaluated at once: instances are numbered per (scene, label) so the
    per-scene means, hinge terms and pairwise centre distances come out of segmented reductions instead of a Python
    loop with a torch.unique + cdist per scene."""

    def __init__(self, classes=20, ignore_label=-100, joint_training_epoch=120, supervise_sp_offset=True,
                 supervise_instance_size=True, semantic_dice=True):
        super().__init__()
        self.classes, self.ignore_label = classes, ignore_label
        self.joint_training_epoch = joint_training_epoch
        self.supervise_sp_offset, self.supervise_instance_size = supervise_sp_offset, supervise_instance_size
        self.semantic_dice = semantic_dice
        self.dim, self.delta_v, self.delta_d = 7, 0.1, 1.5
        self.param_var, self.param_dist, self.param_reg = 1.0, 1.0, 0.001

    def forward(self, inp, epoch):
        out = {}
        semantic_labels, instance_labels = inp['point_labels']
        scores = inp["semantic_scores"]
        loss = ce_dice_loss(scores, semantic_labels, self.ignore_label, self.semantic_dice)      # :56-63
        out["semantic_loss"] = loss
        if epoch > self.joint_training_epoch:
            sp_sem, sp_ins = inp['superpoint_labels']
            valid = (sp_ins != self.ignore_label) & (sp_sem != self.ignore_label)                # :69
            nvalid = valid.sum()
            out["superpoint_semantic_loss"] = ce_dice_loss(inp['sp_semantic'], sp_sem, self.ignore_label, False)   # :72-74
            loss = loss + out["superpoint_semantic_loss"]
            if self.supervise_sp_offset:                                                          # :77-93
                pred, gt = inp['sp_offset_vector']
                dist_ = (pred - gt).abs().sum(-1)
                out["offset_norm_loss"] = (dist_ * valid).sum() / (nvalid + 1e-6)
                gt_n = gt / (gt.norm(p=2, dim=1, keepdim=True) + 1e-8)
                pr_n = pred / (pred.norm(p=2, dim=1, keepdim=True) + 1e-8)
                out["offset_dir_loss"] = (-(gt_n * pr_n).sum(-1) * valid).sum() / (nvalid + 1e-6)
                loss = loss + out["offset_norm_loss"] + out["offset_dir_loss"]
            feats, sp_batch_offsets = inp['sp_discriminative_features']
            out["superpoint_discriminative_loss"] = self.discriminative(feats, sp_ins, valid, sp_batch_offsets)
            loss = loss + out["superpoint_discriminative_loss"]
            if self.supervise_instance_size:                                                      # :117-127
                po, go = inp['sp_occupancy']
                ps, gs = inp['sp_instance_size']
                out["occupancy_loss"] = ((po - go).abs() * valid).sum() / nvalid
                out["instance_size_loss"] = ((ps - gs).abs() * valid).sum() / nvalid
                loss = loss + out["occupancy_loss"] + out["instance_size_loss"]
        return loss, out

    def discriminative(self, feats, labels, valid, sp_batch_offsets):
        """mean over scenes of l_var + l_dist + 0.001 l_reg (:97-113, 157-230)."""
        S = feats.shape[0]
        B = sp_batch_offsets.numel() - 1
        dev = feats.device
        off = sp_batch_offsets.to(dev).long()
        scene = torch.bucketize(torch.arange(S, device=dev), off[1:], right=True)                # scene of every row
        big = 1 << 24                                    # instance ids of a batch stay far below (scannetv2_dataset.py:386)
        key = scene * big + labels.clamp(min=0)
        key = torch.where(valid, key, key.new_full((), -1))
        # a sentinel -1 in front guarantees that bucket 0 is the bucket of the invalid rows (no host round trip to ask)
        uniq, inv, counts = torch.unique(torch.cat([key.new_full((1,), -1), key]), return_inverse=True, return_counts=True)
        uniq, counts, inv = uniq[1:], counts[1:], inv[1:] - 1
        I = uniq.numel()
        w = valid.to(feats.dtype)
        idx = inv.clamp(min=0)
        cnt = counts.to(feats.dtype)
        mu = torch.zeros((I, self.dim), dtype=feats.dtype, device=dev).index_add_(0, idx, feats * w.unsqueeze(1)) / cnt.unsqueeze(1)
        inst_scene = torch.div(uniq, big, rounding_mode="floor")
        n_inst = torch.zeros((B,), dtype=feats.dtype, device=dev).index_add_(0, inst_scene, torch.ones_like(cnt))
        d = (feats - mu[idx]).norm(p=2, dim=1)
        d = torch.clamp(d - self.delta_v, min=0.0) ** 2 * w
        l_var_i = torch.zeros((I,), dtype=feats.dtype, device=dev).index_add_(0, idx, d) / cnt
        l_var = torch.zeros((B,), dtype=feats.dtype, device=dev).index_add_(0, inst_scene, l_var_i) / n_inst
        pd = torch.cdist(mu, mu, p=1)
        same_scene = inst_scene.unsqueeze(0) == inst_scene.unsqueeze(1)
        hinge = torch.clamp(2.0 * self.delta_d - pd, min=0.0) ** 2
        hinge = hinge * (same_scene & ~torch.eye(I, dtype=torch.bool, device=dev))
        l_dist = torch.zeros((B,), dtype=feats.dtype, device=dev).index_add_(0, inst_scene, hinge.sum(1))
        l_dist = torch.where(n_inst > 1, l_dist / (n_inst * (n_inst - 1)).clamp(min=1), torch.zeros_like(l_dist))
        l_reg = torch.zeros((B,), dtype=feats.dtype, device=dev).index_add_(0, inst_scene, mu.norm(p=2, dim=1))
        per_scene = self.param_var * l_var + self.param_dist * l_dist + self.param_reg * l_reg
        return per_scene.mean()


# ---------------------------------------------------------------------------------------------------------
# optimizer + step
# ---------------------------------------------------------------------------------------------------------
class FlatAdamW:
    """torch.optim.AdamW(lr, weight_decay) (train_scannetv2.py:93-94) over ONE flat parameter buffer: parameters are
    re-pointed into `flat_p`, gradients accumulate in the matching flat bucket (dist.GradBucket), the update is one
    kernel.  `clamp_params` (model.ecc.parameters()) are clamped to [-1, 1] first (:246-249)."""

    def __init__(self, model, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, clamp_module=None):
        from .dist import GradBucket
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty((n,), dtype=torch.float32, device=dev)
        off = 0
        self.clamp = (0, 0)
        clamp_ids = {id(p) for p in clamp_module.parameters()} if clamp_module is not None else set()
        cb = ce = None
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            if id(p) in clamp_ids:
                cb = off if cb is None else cb
                assert ce is None or ce == off, "clamped parameters must be contiguous in model.parameters() order"
                ce = off + k
            off += k
        if cb is not None:
            self.clamp = (cb, ce)
        W.invalidate_caches()                      # parameter storage moved
        self.bucket = GradBucket(self.params)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.lr, self.wd, self.betas, self.eps, self.t = lr, weight_decay, betas, eps, 0

    def zero_grad(self):
        self.bucket.zero()

    def step(self, grad_scale=1.0):
        self.t += 1
        lib().call("wsis_adamw_step", _ptr(self.flat_p), _ptr(self.bucket.flat), _ptr(self.m), _ptr(self.v),
                   self.flat_p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                   float(self.wd), self.t, float(grad_scale), self.clamp[0], self.clamp[1], _stream())
        W.invalidate_caches()                      # in-place parameter update: packed weight images are stale


class TrainStep:
    """One data-parallel training step of the reference's loop body (train_scannetv2.py:149-252)."""

    def __init__(self, model, epoch=121, lr=1e-3, weight_decay=1e-4, loss=None):
        self.model = model.train()
        self.loss = loss or MultiTaskLoss(classes=model.classes)
        self.opt = FlatAdamW(model, lr=lr, weight_decay=weight_decay, clamp_module=model.ecc)
        self.epoch = epoch

    def __call__(self, dbatch, optimize=True):
        from . import pipeline
        self.opt.zero_grad()
        with torch.enable_grad():
            ret, aux = pipeline.forward_batch(self.model, dbatch)
            loss, parts = self.loss(loss_inputs(ret, dbatch), self.epoch)
            loss.backward()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.opt.bucket.flat)          # ONE collective per step (NCCL over NVLink)
        if optimize:
            self.opt.step(grad_scale=1.0 / world)
        return loss.detach(), parts


def loss_inputs(ret, dbatch):
    """train_scannetv2.py:211-231."""
    return {
        'point_labels': (dbatch["semantic_labels"], dbatch["instance_labels"]),
        "semantic_scores": ret["semantic_scores"],
        'superpoint_labels': (dbatch["superpoint_semantic_labels"], dbatch["superpoint_instance_labels"]),
        'sp_semantic': ret['sp_semantic_scores'],
        'sp_offset_vector': (ret['pred_sp_offset_vectors'], dbatch["superpoint_offset_vector"]),
        'sp_occupancy': (ret['pred_sp_occupancy'], dbatch["superpoint_instance_voxel_num"]),
        'sp_instance_size': (ret['pred_sp_ins_size'], dbatch["superpoint_instance_size"]),
        'sp_discriminative_features': (ret['sp_discriminative_feats'], dbatch["sp_batch_offsets"]),
    }


__all__ = ["TrainStep", "FlatAdamW", "MultiTaskLoss", "bn_relu_conv", "batch_norm_train", "run_sequential", "ce_dice_loss",
           "gather_rows", "segment_mean", "loss_inputs"]
