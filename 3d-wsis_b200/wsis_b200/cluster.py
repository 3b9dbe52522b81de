"""Instance clustering on the superpoint graph: the last stage between the network outputs and the final instance
masks.

Mirrors `clustering_in_graph` of the reference (test_scannetv2.py:281-455) decision for decision -- breadth-first
merging of same-class neighbouring superpoints whose predicted instance centres are closer than a quarter of the
seed's predicted instance size, a voxel-occupancy test that splits the groups into primary instances and fragments,
and the absorption of every fragment by the nearest primary instance of its class -- so that the masks are identical
on identical inputs.  What changes is the cost: the reference keeps an N-point boolean mask per superpoint and per
group (O(S*N) memory traffic, ~3000 x 150k per scene); here everything is done on per-superpoint aggregates
(point count, centre, voxel keys in one stable sort by superpoint) and the N-point masks are produced once, at the
end, from a superpoint -> instance table.

`clustering_in_graph` is the host restatement (numpy; what the golden vectors of the reference's own function pin);
`clustering_in_graph_device` runs the same algorithm on the GPU (csrc/cluster.cu: one warp walks the graph, the
point-sized work is parallel kernels) and returns device tensors.  `neighbors[s]` is the list the reference gets from
`graph.neighbors(vertex=s, mode='all')` (igraph; ascending vertex ids).
"""
import collections
from math import sqrt

import numpy as np

# test_scannetv2.py:289-290
SEMANTIC_IND2LABEL = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])
INSTANCE_VALID_LABELS = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])


def neighbors_from_edges(edges, num_superpoints):
    """Adjacency lists with igraph's `neighbors(mode='all')` order (ascending ids) from an int[E,2] edge list."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    both = np.concatenate([edges, edges[:, ::-1]])
    both = np.unique(both, axis=0)                      # sorted by (vertex, neighbour)
    starts = np.searchsorted(both[:, 0], np.arange(num_superpoints + 1))
    return [both[starts[s]:starts[s + 1], 1].tolist() for s in range(num_superpoints)]


class _SuperpointTable(object):
    """Per-superpoint aggregates the reference recomputes from N-point masks (get_superpoint_feature, :295-309)."""

    def __init__(self, xyz_origin, superpoint, pred_sp_offset_vectors, voxel_scale):
        self.order = np.argsort(superpoint, kind="stable")     # points of a superpoint in their original order
        sp_sorted = superpoint[self.order]
        S = int(superpoint.max()) + 1
        self.start = np.searchsorted(sp_sorted, np.arange(S + 1))
        self.count = np.diff(self.start)
        xyz_sorted = xyz_origin[self.order]
        centre = np.zeros((S, 3), dtype=xyz_origin.dtype)
        for s in range(S):                                     # the same row-sequential mean as xyz[mask].mean(0)
            if self.count[s]:
                centre[s] = xyz_sorted[self.start[s]:self.start[s + 1]].mean(0)
        self.instance_centre = centre + pred_sp_offset_vectors  # :305
        # voxel key of every point: (xyz * 50) truncated like torch .long() (:373-375), packed into one int64
        v = (xyz_sorted * voxel_scale).astype(np.int64)
        v -= v.min(0) if len(v) else 0
        self.vkey = (v[:, 0] << 42) | (v[:, 1] << 21) | v[:, 2]

    def group_voxels(self, group):
        """Number of distinct voxels the points of the superpoints in `group` fall into (= voxel_locs.shape[0] of
        pointgroup_ops.voxelization_idx on the group's points, :376-377)."""
        keys = np.concatenate([self.vkey[self.start[s]:self.start[s + 1]] for s in group])
        return int(np.unique(keys).shape[0])


def clustering_in_graph(xyz_origin, superpoint, neighbors, sp_semantic_pred, pred_sp_offset_vectors, pred_sp_occupancy,
                        pred_sp_ins_size, semantic_ind2label=SEMANTIC_IND2LABEL,
                        valid_labels=INSTANCE_VALID_LABELS, voxel_scale=50, dense=True):
    """-> (conf float[I], label_id int[I], masks int[I, N]); same contract as test_scannetv2.py:281-455.
    dense=False returns the point -> instance table int64[N] (-1 = none) instead of the I x N masks."""
    xyz_origin = np.asarray(xyz_origin)
    superpoint = np.asarray(superpoint)
    assert len(xyz_origin) == len(superpoint)
    sp_ids = np.unique(superpoint)
    assert len(sp_ids) == (superpoint.max() + 1) == len(sp_semantic_pred) == len(pred_sp_offset_vectors)
    tab = _SuperpointTable(xyz_origin, superpoint, pred_sp_offset_vectors, voxel_scale)
    centre, count = tab.instance_centre, tab.count
    visited = {int(s): False for s in sp_ids}
    valid = set(int(x) for x in valid_labels)

    def bfs(seed):                                             # :313-345
        visited[seed] = True
        queue = collections.deque([seed])
        group = set([seed])
        label = sp_semantic_pred[seed]
        while queue:
            cur = queue.popleft()
            for nb in neighbors[cur]:
                if sp_semantic_pred[nb] == label and not visited[nb]:
                    if np.linalg.norm(centre[cur] - centre[nb], ord=2) < 0.25 * pred_sp_ins_size[seed]:
                        group.add(nb)
                        visited[nb] = True
                        queue.append(nb)
        return list(group)

    def group_occupancy(group):                                # :352-356
        return np.exp(pred_sp_occupancy[np.array(group)]).mean()

    def group_centre(group):                                   # :359-367
        c = np.zeros(3)
        n = 0
        for s in group:
            c += centre[s] * count[s]
            n += count[s]
        return c / n

    def group_size(group):                                     # :369-371
        return np.mean(pred_sp_ins_size[np.array(group)])

    primaries, fragments = [], []
    for seed in sp_ids.tolist():                               # :375-414
        label = sp_semantic_pred[seed]
        if (int(semantic_ind2label[label]) not in valid) or visited[seed]:
            continue
        group = bfs(seed)
        occ = group_occupancy(group)
        n_points = count[np.array(group)].sum()             # np.int64, like group_mask.sum()
        if tab.group_voxels(group) < 0.3 * occ:
            fragments.append({"classLabel": label, "instance_center": group_centre(group), "group_sp_list": group,
                              "group_n": n_points})
        else:
            r_set = max(0.01 * sqrt(n_points), 0.02 * sqrt(occ), group_size(group))
            primaries.append({"classLabel": label, "instance_center": group_centre(group), "r_set": r_set,
                              "group_sp_list": group, "group_n": n_points})

    if primaries:
        for frag in fragments:                                 # :418-446
            index, dis_min = -1, float("inf")
            for i, prim in enumerate(primaries):
                dis = np.linalg.norm(frag["instance_center"] - prim["instance_center"], ord=2)
                if frag["classLabel"] == prim["classLabel"] and dis < dis_min:
                    index, dis_min = i, dis
            closest = primaries[index]
            if dis_min < closest["r_set"]:
                merged = frag["group_sp_list"] + closest["group_sp_list"]
                n_points = frag["group_n"] + closest["group_n"]   # the masks are disjoint
                closest["r_set"] = max(0.02 * sqrt(group_occupancy(merged)), 0.01 * sqrt(n_points), closest["r_set"],
                                       group_size(merged))
                closest["instance_center"] = group_centre(merged)
                closest["group_n"] = n_points
                closest["group_sp_list"] += frag["group_sp_list"]

    conf, label_id = [], []
    inst_of_sp = np.full(int(superpoint.max()) + 1, -1, dtype=np.int64)
    for i, prim in enumerate(primaries):                       # :449-457
        conf.append(min(prim["group_n"] / group_occupancy(prim["group_sp_list"]), 1))
        label_id.append(semantic_ind2label[prim["classLabel"]])
        inst_of_sp[np.array(prim["group_sp_list"])] = i
    point_inst = inst_of_sp[superpoint]
    if not dense:
        return np.array(conf), np.array(label_id), point_inst
    masks = (point_inst[None, :] == np.arange(len(primaries))[:, None]).astype(int) if primaries \
        else np.zeros((0, len(superpoint)), dtype=int)
    return np.array(conf), np.array(label_id), masks


def neighbors_csr_device(edges, num_superpoints):
    """(nbr_off int32[S+1], nbr int32[...]) on the device from an int64[E,2] CUDA edge list: both directions, duplicates
    removed, ascending neighbour id per vertex (= neighbors_from_edges)."""
    import torch
    both = torch.cat([edges, edges.flip(1)]).long()
    key = torch.unique(both[:, 0] * num_superpoints + both[:, 1])          # sorted by (vertex, neighbour)
    src, nbr = torch.div(key, num_superpoints, rounding_mode="floor"), key % num_superpoints
    off = torch.zeros(num_superpoints + 1, dtype=torch.int64, device=edges.device)
    off[1:] = torch.cumsum(torch.bincount(src, minlength=num_superpoints), 0)
    return off.int().contiguous(), nbr.int().contiguous()


def clustering_in_graph_device(xyz_origin, superpoint, nbr_csr, sp_semantic_pred, pred_sp_offset_vectors, pred_sp_occupancy,
                               pred_sp_ins_size, num_superpoints=None, semantic_ind2label=SEMANTIC_IND2LABEL,
                               valid_labels=INSTANCE_VALID_LABELS, voxel_scale=50, sp_index=None):
    """Device version of clustering_in_graph for one scene.  CUDA tensors in: xyz_origin f32[N,3], superpoint int64[N],
    nbr_csr = neighbors_csr_device(...), sp_semantic_pred int[S], offsets f32[S,3], occupancy f32[S], size f32[S].
    Returns (conf float64[I], label_id int32[I], point_inst int32[N], inst_of_sp int32[S]) as CUDA tensors; the
    reference's dense masks are `dense_masks(point_inst, I)`.  One host sync (the number of instances)."""
    import ctypes

    import torch

    from . import ops as W
    from ._lib import lib
    xyz = xyz_origin.contiguous().float()
    superpoint = superpoint.contiguous().long()
    N = xyz.shape[0]
    S = int(num_superpoints) if num_superpoints is not None else int(sp_semantic_pred.shape[0])
    dev = xyz.device
    seg = sp_index if sp_index is not None else W.SegmentIndex(superpoint, S)
    centre = (W.segment_reduce(xyz, seg, "mean") + pred_sp_offset_vectors.float()).contiguous()       # :295-305
    count = (seg.offsets[1:] - seg.offsets[:-1]).int().contiguous()
    n_class = len(semantic_ind2label)
    valid = torch.tensor([1 if int(x) in set(int(v) for v in valid_labels) else 0 for x in semantic_ind2label],
                         dtype=torch.int32, device=dev)
    ind2label = torch.tensor([int(x) for x in semantic_ind2label], dtype=torch.int32, device=dev)
    sem = sp_semantic_pred.int().contiguous()
    occ, size = pred_sp_occupancy.float().contiguous(), pred_sp_ins_size.float().contiguous()
    nbr_off, nbr = nbr_csr
    ws = torch.empty(int(lib().value("wsis_cluster_ws_bytes", N, S)) + 256, dtype=torch.uint8, device=dev)
    conf = torch.empty((S,), dtype=torch.float64, device=dev)
    label_id = torch.empty((S,), dtype=torch.int32, device=dev)
    inst_of_sp = torch.empty((S,), dtype=torch.int32, device=dev)
    point_inst = torch.empty((N,), dtype=torch.int32, device=dev)
    n_inst = torch.zeros((1,), dtype=torch.int32, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    lib().call("wsis_cluster", p(xyz), p(superpoint), N, S, p(nbr_off), p(nbr), p(sem), p(centre), p(count), p(occ), p(size),
               p(valid), p(ind2label), n_class, float(voxel_scale), p(ws), p(conf), p(label_id), p(inst_of_sp),
               p(point_inst), p(n_inst), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    n = int(n_inst.item())
    return conf[:n], label_id[:n], point_inst, inst_of_sp


def dense_masks(point_inst, n_instances):
    """int[I, N] masks of the reference (test_scannetv2.py:449-457) from the point -> instance table."""
    import torch
    return (point_inst.unsqueeze(0) == torch.arange(n_instances, device=point_inst.device, dtype=point_inst.dtype).unsqueeze(1)).long()
