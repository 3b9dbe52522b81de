"""Scene-level hot path: the step body of the reference's drivers, from a collated host batch to the network
outputs and (optionally) the propagated pseudo labels.

    forward:      train_scannetv2.py:149-198 / test_scannetv2.py:140-190
                  (.cuda() copies -> superpoint centres -> voxelization -> SparseConvTensor -> Network)
    propagation:  train_scannetv2.py:554-575 -> modules/datasets/scannetv2_dataset.py:664-735
"""
from types import SimpleNamespace

import torch

import pointgroup_ops
import spconv

from . import ops as W
from .model import GraphInfo, Network

DEFAULT_MODEL_CFG = dict(input_channel=3, use_coords=True, blocks=5, block_reps=2, media=32, classes=20,
                         fix_module="[]")  # config/ScanNet_v2_3D_WSIS.yaml:37-47

# tensors the forward needs on the device (what train_scannetv2.py:149-172 moves with .cuda())
_DEVICE_KEYS = ("locs", "locs_float", "feats", "superpoint", "edge_u_list", "edge_v_list", "ecc_edge_index",
                "ecc_edgefeats", "seed_label",
                # training labels (train_scannetv2.py:155-167)
                "semantic_labels", "instance_labels", "superpoint_semantic_labels", "superpoint_instance_labels",
                "superpoint_offset_vector", "superpoint_instance_voxel_num", "superpoint_instance_size")


def build_network(cfg=None, seed=123, device="cuda"):
    """Random-init network of the reference architecture under torch.manual_seed(seed) (config/*.yaml:3)."""
    cfg = dict(DEFAULT_MODEL_CFG, **(cfg or {}))
    torch.manual_seed(seed)
    net = Network(SimpleNamespace(**cfg))
    return net.to(device)


def pin_batch(batch):
    return {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}


def to_device(batch, device="cuda", non_blocking=True):
    """Host -> device copy of one collated batch; returns (device batch, bytes copied)."""
    out, nbytes = dict(batch), 0
    for k in _DEVICE_KEYS:
        if k in batch:
            out[k] = batch[k].to(device, non_blocking=non_blocking)
            nbytes += batch[k].numel() * batch[k].element_size()
    return out, nbytes


class BatchStream(object):
    """Iterates over collated HOST batches (pinned: pin_batch) and yields (device batch, bytes copied): the host -> device
    copy of batch i+1 is issued on a copy stream while batch i is being computed on the current stream, the way the
    reference's DataLoader(pin_memory=True) + .cuda(non_blocking) loop overlaps them (train_scannetv2.py:149-172).
    Every batch is still copied exactly once; nothing is cached across iterations.

    The device side is `depth + 1` sets of persistent staging buffers (grown on demand), so the two streams never meet
    in the caching allocator: set k is overwritten only after the compute stream has finished the batch that used it
    (an event recorded when the consumer asks for the next batch)."""

    def __init__(self, host_batches, device="cuda", depth=1, copy_stream=None, staging=None, prepare=False):
        self.host_batches, self.device, self.depth = host_batches, device, max(int(depth), 1)
        # high priority: the loader's small kernels (geometry build) must get SMs as the persistent conv CTAs of the
        # compute stream retire, not after its whole queue has drained
        self.copy_stream = copy_stream if copy_stream is not None else torch.cuda.Stream(priority=-1)
        # a long-lived loader passes its staging sets back in (BatchStream(..., staging=prev.staging))
        self.staging = staging if staging is not None else [dict(bufs={}, done=None) for _ in range(self.depth + 1)]
        # prepare=True: the coordinate-only part of the step (prepare_geometry: voxelization maps, rulebooks, tile
        # records) is built on the copy stream as well, i.e. under the previous batch's feature compute, and attached to
        # the batch as "_geometry" (forward_batch consumes it)
        self.prepare = prepare
        self._keep = []          # (event, objects): side-stream allocations stay referenced until their consumer is done

    def _issue(self, batch, slot):
        st = self.staging[slot]
        out, nbytes = dict(batch), 0
        with torch.cuda.stream(self.copy_stream):
            if st["done"] is not None:
                self.copy_stream.wait_event(st["done"])          # the batch that used this set has been computed
            for k in _DEVICE_KEYS:
                if k not in batch:
                    continue
                h = batch[k]
                buf = st["bufs"].get(k)
                if buf is None or buf.numel() < h.numel() or buf.dtype != h.dtype:
                    buf = torch.empty((int(h.numel() * 1.25) + 16,), dtype=h.dtype, device=self.device)
                    st["bufs"][k] = buf
                view = buf[:h.numel()].view(h.shape)
                view.copy_(h, non_blocking=True)
                out[k] = view
                nbytes += h.numel() * h.element_size()
            if self.prepare:
                out["_geometry"] = prepare_geometry(out)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return out, nbytes, ev, slot

    def __iter__(self):
        it = iter(self.host_batches)
        n_sets, slot = len(self.staging), 0
        first = next(it, None)
        cur = self._issue(first, slot) if first is not None else None      # nothing to overlap the first batch with
        while cur is not None:
            db, nb, ev, used = cur
            torch.cuda.current_stream().wait_event(ev)
            yield db, nb
            # the consumer is back: everything that reads `db` is queued on the compute stream
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())
            self.staging[used]["done"] = done
            self._keep.append((done, db))                       # tensors allocated on the copy stream: alive until `done`
            self._keep = [(e, o) for e, o in self._keep if not e.query()]
            nxt = next(it, None)
            slot = (slot + 1) % n_sets
            cur = self._issue(nxt, slot) if nxt is not None else None      # H2D (+ geometry) under the compute of `db`


class ResultFetcher(object):
    """Asynchronous device -> host read of a step's results into pinned buffers (stream-ordered, no host sync per
    step); `wait()` blocks until everything fetched so far has landed."""

    def __init__(self, keys=("edge_affinity", "sp_semantic_scores", "sp_discriminative_feats", "pred_sp_offset_vectors")):
        self.keys, self.bufs, self.last = keys, {}, None

    def fetch(self, ret):
        out, nbytes = {}, 0
        for k in self.keys:
            t = ret[k]
            buf = self.bufs.get(k)
            if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                buf = torch.empty((int(t.numel() * 1.25) + 16,), dtype=t.dtype, pin_memory=True)
                self.bufs[k] = buf
            view = buf[:t.numel()].view(t.shape)
            view.copy_(t, non_blocking=True)
            out[k] = view
            nbytes += t.numel() * t.element_size()
        self.last = torch.cuda.Event()
        self.last.record()
        return out, nbytes

    def wait(self):
        if self.last is not None:
            self.last.synchronize()


def unet_rulebook_keys(blocks=5):
    """The `indice_key`s the U-Net's convolutions look their rulebooks up by (sparse_unet3d.py:229-350, backbone_3D_WSIS.py:
    46-50): one submanifold rulebook per level, one strided rulebook per down / up pair."""
    return ["subm%d" % level for level in range(1, blocks + 1)] + ["spconv%d" % level for level in range(1, blocks)]


def prepare_geometry(dbatch, mode=4, blocks=5):
    """Everything of a step that depends on COORDINATES only (inference): voxelization maps, the superpoint / edge
    segment indices, and the nine rulebooks of the U-Net with their tile records (conv.py:140-152 builds them lazily
    inside the first conv that needs them; sparse_unet3d.py:229-350 fixes which).  No feature is read, so a loader can
    run this for batch i+1 on its side stream while batch i's features are being computed (BatchStream(prepare=True)).
    Returns the dict forward_batch(..., geometry=...) consumes."""
    S = dbatch["num_superpoints"]
    bs = dbatch["batch_size"]
    voxel_locs, p2v_map, v2p_map = pointgroup_ops.voxelization_idx(dbatch["locs"], bs, mode)
    geo = {"voxel_locs": voxel_locs, "p2v_map": p2v_map, "v2p_map": v2p_map,
           "sp_index": W.SegmentIndex(dbatch["superpoint"], S), "edge_index_u": W.SegmentIndex(dbatch["edge_u_list"], S)}
    coords, shape = voxel_locs.int(), list(dbatch["spatial_shape"])
    geo["coords"] = coords
    indice_dict = {}
    for level in range(1, blocks + 1):
        rb = W.rulebook_subm(coords, shape, 3, 1, bs)
        rb.tiles_out()
        lazy = spconv.ops.LazyPairs(rb)
        indice_dict["subm%d" % level] = (coords, coords, lazy, lazy, shape)
        if level < blocks:
            rbc, oshape = W.rulebook_conv(coords, shape, 2, 2, 0, 1, bs)
            rbc.tiles_out()                                   # strided conv (down)
            rbc.tiles_in()                                    # inverse conv (up)
            lazy = spconv.ops.LazyPairs(rbc)
            indice_dict["spconv%d" % level] = (rbc.out_coords, coords, lazy, lazy, shape)
            coords, shape = rbc.out_coords, oshape
    assert sorted(indice_dict) == sorted(unet_rulebook_keys(blocks))
    geo["indice_dict"] = indice_dict
    return geo


def forward_batch(model, dbatch, use_coords=True, mode=4, keep_unet_features=False, geometry=None):
    """Voxelization + UNet + pooling + affinity for one device-resident batch.  Returns (ret dict, aux dict).
    `geometry` = prepare_geometry(dbatch) when a loader has already built the coordinate-only part (inference)."""
    locs = dbatch["locs"]
    S = dbatch["num_superpoints"]
    geometry = geometry if geometry is not None else dbatch.get("_geometry")
    if geometry is not None and (torch.is_grad_enabled() or model.training):
        geometry = None                                       # the training path keeps the reference-format pairs
    if geometry is not None:
        voxel_locs, p2v_map, v2p_map = geometry["voxel_locs"], geometry["p2v_map"], geometry["v2p_map"]
    else:
        voxel_locs, p2v_map, v2p_map = pointgroup_ops.voxelization_idx(locs, dbatch["batch_size"], mode)
    coords_float = dbatch["locs_float"]
    superpoint = dbatch["superpoint"]
    sp_index = geometry["sp_index"] if geometry is not None else W.SegmentIndex(superpoint, S)
    centers = W.segment_reduce(coords_float, sp_index, "mean")                       # train_scannetv2.py:177
    feats = torch.cat((dbatch["feats"], coords_float), 1) if use_coords else dbatch["feats"]
    voxel_feats = pointgroup_ops.voxelization(feats, v2p_map, mode)                 # :189
    input_ = spconv.SparseConvTensor(voxel_feats, geometry["coords"] if geometry is not None else voxel_locs.int(),
                                     dbatch["spatial_shape"], dbatch["batch_size"])
    if geometry is not None:
        input_.indice_dict = dict(geometry["indice_dict"])    # every conv finds its rulebook (conv.py:140-147)
    eindex = geometry["edge_index_u"] if geometry is not None else W.SegmentIndex(dbatch["edge_u_list"], S)
    extra = {"superpoint": superpoint, "GIs": [GraphInfo(dbatch["ecc_edge_index"], dbatch["ecc_edgefeats"])],
             "edge_u_list": dbatch["edge_u_list"], "edge_v_list": dbatch["edge_v_list"],
             "superpoint_cenetr_xyz": centers, "sp_index": sp_index, "edge_index_u": eindex, "num_superpoints": S,
             "keep_unet_features": keep_unet_features}
    if torch.is_grad_enabled() or model.training:
        ret = model(input_, p2v_map, extra)
    else:
        with spconv.ops.lazy_pairs():  # inference: nothing reads the reference-format pair tensors
            ret = model(input_, p2v_map, extra)
    aux = {"voxel_locs": voxel_locs, "p2v_map": p2v_map, "v2p_map": v2p_map, "centers": centers, "input": input_,
           "edge_index_u": eindex, "unet_features": extra.get("unet_features")}
    return ret, aux


def propagate_labels(ret, dbatch, aux, iterations, class_num=20):
    """Random-walk pseudo labels for a batch of ONE scene (the reference propagates with batch_size=1,
    train_scannetv2.py:498).  Returns (pseudo int32[S], score float64[S])."""
    prob = torch.softmax(ret['sp_semantic_scores'], dim=-1)                          # train_scannetv2.py:555-557
    conf, pred = prob.max(1)
    vseg = W.SegmentIndex(dbatch["edge_v_list"], dbatch["num_superpoints"])
    return W.random_walk(dbatch["edge_u_list"], dbatch["edge_v_list"], ret['edge_affinity'], dbatch["seed_label"],
                         pred, conf, class_num, iterations, useg=aux["edge_index_u"], vseg=vseg)
