"""Pseudo-label bookkeeping on the superpoint graph between training stages: the 1-hop label
extension, the 1-hop propagation and the whole-scene propagation of modules/datasets/scannetv2_dataset.py:779-967.

The reference keeps the labels as igraph vertex attributes and recomputes every superpoint centre from an N-point
boolean mask each time it is needed (O(S*N) per call).  Here the state is three arrays over the superpoints
(`semantic`, `instance`, `offset`), the geometry is a one-off table of per-superpoint aggregates (count, coordinate
sum, centre -- the same row-sequential float32 arithmetic as `xyz[mask].mean(0)` / `.sum(0)`), and the decisions are
taken in the reference's order, so the results are identical.  Host control logic; no kernel.
"""
import collections

import numpy as np

UNLABELED = -100


class SuperpointGeometry(object):
    """count[S], xyz_sum[S,3] (= xyz[mask].sum(0)), centre[S,3] (= xyz[mask].mean(0)), in the input dtype."""

    def __init__(self, xyz_origin, superpoint):
        xyz_origin, superpoint = np.asarray(xyz_origin), np.asarray(superpoint)
        order = np.argsort(superpoint, kind="stable")          # points of a superpoint in their original order
        S = int(superpoint.max()) + 1
        start = np.searchsorted(superpoint[order], np.arange(S + 1))
        xs = xyz_origin[order]
        self.count = np.diff(start)
        self.xyz_sum = np.zeros((S, 3), dtype=xyz_origin.dtype)
        self.centre = np.zeros((S, 3), dtype=xyz_origin.dtype)
        for s in range(S):
            if self.count[s]:
                seg = xs[start[s]:start[s + 1]]
                self.xyz_sum[s] = seg.sum(0)
                self.centre[s] = seg.mean(0)


class SuperpointLabels(object):
    """semantic int[S], instance int[S] (-100 = unlabeled), offset float[S,3] (superpoint -> instance centre)."""

    def __init__(self, semantic, instance, offset):
        self.semantic = np.array(semantic)
        self.instance = np.array(instance)
        self.offset = np.array(offset, dtype=np.float64)

    def copy(self):
        return SuperpointLabels(self.semantic, self.instance, self.offset)

    def labeled(self, s):
        return self.semantic[s] != UNLABELED and self.instance[s] != UNLABELED


def _one_hop(labels, geo, neighbors, sp_semantic_pred, accept):
    out = labels.copy()
    for ind in range(len(labels.semantic)):
        if not labels.labeled(ind):
            continue
        for nb in neighbors[ind]:
            if (sp_semantic_pred[nb] == labels.semantic[ind]) and accept(nb) \
                    and labels.semantic[nb] == UNLABELED and labels.instance[nb] == UNLABELED:
                out.semantic[nb] = labels.semantic[ind]
                out.instance[nb] = labels.instance[ind]
                instance_centre = geo.centre[ind] + labels.offset[ind]
                out.offset[nb] = instance_centre - geo.centre[nb]
    return out


def extend_label_to_neighbor(labels, geo, neighbors, sp_semantic_value, sp_semantic_pred, min_confidence=0.8):
    """scannetv2_dataset.py:779-812: every labeled superpoint hands its labels to the unlabeled graph neighbours whose
    predicted class agrees with confidence > 0.8; the neighbour's offset points at the donor's instance centre."""
    return _one_hop(labels, geo, neighbors, sp_semantic_pred, lambda nb: sp_semantic_value[nb] > min_confidence)


def propagate_label_to_neighbor(labels, geo, neighbors, sp_semantic_value, sp_semantic_pred):
    """scannetv2_dataset.py:828-853: the same hop from the current weak labels, without the confidence test."""
    return _one_hop(labels, geo, neighbors, sp_semantic_pred, lambda nb: True)


def edge_same_instance(edges, labels):
    """is1ins per edge (:814-823): 0 = an endpoint is unlabeled, -1 = same instance, 1 = different instances."""
    edges = np.asarray(edges).reshape(-1, 2)
    a, b = labels.instance[edges[:, 0]], labels.instance[edges[:, 1]]
    return np.where((a == UNLABELED) | (b == UNLABELED), 0, np.where(a == b, -1, 1))


def propagate_label_to_whole_scene(labels, geo, sp_semantic_pred, pred_sp_offset_vectors, max_distance=0.9):
    """scannetv2_dataset.py:872-958: every unlabeled superpoint joins the labeled ("prior") superpoint of its predicted
    class whose instance centre is nearest to its own predicted instance centre, if that is within 0.9 m; the members
    of a prior get its labels and an offset to the centroid of all member points."""
    prior = [s for s in range(len(labels.semantic)) if labels.labeled(s)]
    prior_centre = np.array([geo.centre[s] + labels.offset[s] for s in prior])
    prior_instance = np.array([labels.instance[s] for s in prior])
    prior_semantic = np.array([labels.semantic[s] for s in prior])
    out = labels.copy()
    members = collections.defaultdict(set)
    for s in range(len(labels.semantic)):
        if labels.labeled(s):
            continue
        pred_centre = geo.centre[s] + pred_sp_offset_vectors[s]
        if (prior_semantic == sp_semantic_pred[s]).sum() == 0:
            continue
        selected = np.where(prior_semantic == sp_semantic_pred[s])[0]
        dist = np.linalg.norm(prior_centre[selected] - pred_centre, ord=2, axis=1)
        closest = np.argmin(dist)
        if dist[closest] > max_distance:
            continue
        members[selected[closest]].add(s)
    for p, sp_set in members.items():
        sp_list = list(sp_set)
        centroid = np.zeros(3)
        n = 0
        for s in sp_list:
            centroid += geo.xyz_sum[s]
            n += geo.count[s]
        centroid = centroid / n
        for s in sp_list:
            out.semantic[s] = prior_semantic[p]
            out.instance[s] = prior_instance[p]
            out.offset[s] = centroid - geo.centre[s]
    return out
