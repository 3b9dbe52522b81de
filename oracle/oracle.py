"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front end of the CPU oracle (oracle/wsis_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.  Every function cites the reference file:line whose
algorithm the C code restates (citations are into /root/reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """gcc the plain-C restatement (seconds).  Building the checker is not using it."""
    so = os.path.join(_HERE, "libwsis_oracle.so")
    src = os.path.join(_HERE, "wsis_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_rulebook_conv.restype = ctypes.c_int64
        _LIB.orc_voxelization_idx.restype = ctypes.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _triple(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return np.asarray(v, dtype=np.int32).copy()
    return np.asarray([v] * 3, dtype=np.int32)


def rulebook_subm(coords, batch_size, spatial_shape, ksize=3, dilation=1):
    """getIndicePair<3> subm branch, spconv_ops.h:86-102 -> geometry.h:246-293.
    Returns (pairs int32[K,2,N] padded with -1, num int32[K])."""
    coords = _i32(coords)
    N = coords.shape[0]
    ks, dil, shape = _triple(ksize), _triple(dilation), _triple(spatial_shape)
    K = int(ks.prod())
    pairs = np.full((K, 2, N), -1, np.int32)
    num = np.zeros(K, np.int32)
    rc = lib().orc_rulebook_subm(_p(coords, _i32p), ctypes.c_int64(N), ctypes.c_int32(batch_size),
                                 _p(shape, _i32p), _p(ks, _i32p), _p(dil, _i32p), _p(pairs, _i32p), _p(num, _i32p))
    assert rc == 0
    return pairs, num


def conv_output_shape(spatial_shape, ksize, stride, padding, dilation):
    """ops.get_conv_output_size, spconv/ops.py:19-30."""
    s, k, st, p, d = (_triple(x) for x in (spatial_shape, ksize, stride, padding, dilation))
    return [int((s[i] + 2 * p[i] - d[i] * (k[i] - 1) - 1) // st[i] + 1) for i in range(3)]


def rulebook_conv(coords, batch_size, spatial_shape, ksize, stride, padding=0, dilation=1):
    """getIndicePair<3> regular-conv branch on CPU, spconv_ops.h:103-136 -> geometry.h:145-194.
    Returns (out_coords int32[M,4] in first-touch order, pairs int32[K,2,N], num int32[K], out_shape)."""
    coords = _i32(coords)
    N = coords.shape[0]
    ks, st, pad, dil = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    oshape = np.asarray(conv_output_shape(spatial_shape, ks, st, pad, dil), np.int32)
    K = int(ks.prod())
    pairs = np.full((K, 2, N), -1, np.int32)
    num = np.zeros(K, np.int32)
    outc = np.zeros((max(N * K, 1), 4), np.int32)
    M = lib().orc_rulebook_conv(_p(coords, _i32p), ctypes.c_int64(N), ctypes.c_int32(batch_size), _p(oshape, _i32p),
                                _p(ks, _i32p), _p(st, _i32p), _p(pad, _i32p), _p(dil, _i32p),
                                _p(outc, _i32p), _p(pairs, _i32p), _p(num, _i32p))
    assert M >= 0
    return outc[:M].copy(), pairs, num, [int(x) for x in oshape]


def indice_conv(features, filters, pairs, num, n_out, inverse=False):
    """indiceConv<float>, spconv_ops.h:253-349.  filters: [..., Cin, Cout]."""
    features = np.ascontiguousarray(features, np.float32)
    Cin, Cout = filters.shape[-2], filters.shape[-1]
    filt = np.ascontiguousarray(filters, np.float32).reshape(-1, Cin, Cout)
    pairs, num = _i32(pairs), _i32(num)
    K = filt.shape[0]
    assert pairs.shape[0] == K
    out = np.zeros((n_out, Cout), np.float32)
    lib().orc_indice_conv_fwd(_p(features, _f32p), _p(filt, _f32p), _p(pairs, _i32p), _p(num, _i32p),
                              ctypes.c_int64(pairs.shape[2]), K, Cin, Cout, ctypes.c_int64(n_out),
                              int(bool(inverse)), _p(out, _f32p))
    return out


def indice_conv_backward(features, filters, dout, pairs, num, inverse=False):
    """indiceConvBackward<float>, spconv_ops.h:351-433.  Returns (din, dfilters)."""
    features = np.ascontiguousarray(features, np.float32)
    dout = np.ascontiguousarray(dout, np.float32)
    Cin, Cout = filters.shape[-2], filters.shape[-1]
    filt = np.ascontiguousarray(filters, np.float32).reshape(-1, Cin, Cout)
    pairs, num = _i32(pairs), _i32(num)
    K = filt.shape[0]
    din = np.zeros_like(features)
    dfilt = np.zeros_like(filt)
    lib().orc_indice_conv_bwd(_p(features, _f32p), _p(filt, _f32p), _p(dout, _f32p), _p(pairs, _i32p), _p(num, _i32p),
                              ctypes.c_int64(pairs.shape[2]), K, Cin, Cout, ctypes.c_int64(features.shape[0]),
                              int(bool(inverse)), _p(din, _f32p), _p(dfilt, _f32p))
    return din, dfilt.reshape(filters.shape)


def voxelization_idx(coords, batch_size=None, mode=4):
    """pointgroup_ops.voxelization_idx call-site contract (scannetv2_dataset.py:449) -- parity unpinned.
    coords int64[N,4] -> (voxel_locs int64[M,4], p2v int32[N], v2p int32[M,1+maxActive])."""
    coords = np.ascontiguousarray(coords, np.int64)
    N = coords.shape[0]
    ma = ctypes.c_int32(0)
    M = lib().orc_voxelization_idx(_p(coords, _i64p), ctypes.c_int64(N), None, None, None, 0, ctypes.byref(ma))
    locs = np.zeros((M, 4), np.int64)
    p2v = np.zeros(N, np.int32)
    v2p = np.zeros((M, 1 + ma.value), np.int32)
    lib().orc_voxelization_idx(_p(coords, _i64p), ctypes.c_int64(N), _p(locs, _i64p), _p(p2v, _i32p), _p(v2p, _i32p),
                               ctypes.c_int32(1 + ma.value), ctypes.byref(ma))
    return locs, p2v, v2p


def voxelization(feats, v2p, mode=4):
    """pointgroup_ops.voxelization mean (train_scannetv2.py:189) -- parity unpinned."""
    assert mode == 4
    feats = np.ascontiguousarray(feats, np.float32)
    v2p = _i32(v2p)
    M, C = v2p.shape[0], feats.shape[1]
    out = np.zeros((M, C), np.float32)
    lib().orc_voxelization_fwd(_p(feats, _f32p), _p(v2p, _i32p), ctypes.c_int64(M), ctypes.c_int32(v2p.shape[1]), C, _p(out, _f32p))
    return out


def voxelization_backward(dout, v2p, n_points, mode=4):
    dout = np.ascontiguousarray(dout, np.float32)
    v2p = _i32(v2p)
    M, C = dout.shape
    df = np.zeros((n_points, C), np.float32)
    lib().orc_voxelization_bwd(_p(dout, _f32p), _p(v2p, _i32p), ctypes.c_int64(M), ctypes.c_int32(v2p.shape[1]), C,
                               ctypes.c_int64(n_points), _p(df, _f32p))
    return df


_REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2}


def scatter(src, index, reduce="mean", dim_size=None):
    """torch_scatter.scatter(src, index, dim=0, reduce=...) as used at backbone_3D_WSIS.py:188,225,232,244."""
    src = np.ascontiguousarray(src, np.float32)
    squeeze = src.ndim == 1
    src2 = src.reshape(src.shape[0], -1)
    index = np.ascontiguousarray(index, np.int64)
    S = int(index.max()) + 1 if dim_size is None else int(dim_size)
    out = np.zeros((S, src2.shape[1]), np.float32)
    lib().orc_scatter(_p(src2, _f32p), _p(index, _i64p), ctypes.c_int64(src2.shape[0]), src2.shape[1],
                      ctypes.c_int64(S), _REDUCE[reduce], _p(out, _f32p))
    return out[:, 0] if squeeze else out


def edge_attention(q, k, v, ecc, centers, eu, ev, w1, b1, w2, b2):
    """backbone_3D_WSIS.py:209-249.  Returns (edge_affinity f32[E], sp_feat f32[S,D])."""
    q, k, v, ecc, centers = (np.ascontiguousarray(a, np.float32) for a in (q, k, v, ecc, centers))
    w1, b1, w2, b2 = (np.ascontiguousarray(a, np.float32) for a in (w1, b1, w2, b2))
    eu, ev = np.ascontiguousarray(eu, np.int64), np.ascontiguousarray(ev, np.int64)
    S, D = q.shape
    E = eu.shape[0]
    aff = np.zeros(E, np.float32)
    sp = np.zeros((S, D), np.float32)
    lib().orc_edge_attention(_p(q, _f32p), _p(k, _f32p), _p(v, _f32p), _p(ecc, _f32p), _p(centers, _f32p),
                             _p(eu, _i64p), _p(ev, _i64p), ctypes.c_int64(S), ctypes.c_int64(E), D,
                             _p(w1, _f32p), _p(b1, _f32p), _p(w2, _f32p), _p(b2, _f32p), _p(aff, _f32p), _p(sp, _f32p))
    return aff, sp


def ecc_gru_step(h, filters, src, tgt, w_ig, b_ig, w_ih, w_hh, b_ih, b_hh, layernorm=True, eps=1e-5):
    """One step of the edge-conditioned GRU in float64 numpy.
    NNConv with aggr='mean', no root weight, no bias (modules/model/spg_modules.py:97-121: messages x[src]^T W_e,
    averaged at the edge's target; superpoints without an in-edge get 0), then GRUCellEx (:226-253): input gate
    sigmoid(ig(h)) * m, gate pre-activations inp.W_ih^T and h.W_hh^T each normalised by InstanceNorm1d(1) = un-affine
    layer norm with biased variance, reset / update / new gates, h' = n + z (h - n).
    h [S,F]; filters [E,F*F] ([in][out]); src, tgt int[E]; w_ig [F,F], b_ig [F]; w_ih, w_hh [3F,F]; b_ih, b_hh [3F]."""
    h = np.asarray(h, np.float64)
    S, F = h.shape
    W = np.asarray(filters, np.float64).reshape(-1, F, F)
    msg = np.einsum("ei,eio->eo", h[src], W)
    m = np.zeros((S, F))
    np.add.at(m, tgt, msg)
    m /= np.maximum(np.bincount(tgt, minlength=S), 1)[:, None]

    def sig(x):
        return 1.0 / (1.0 + np.exp(-x))

    def ln(x):
        x = x - x.mean(1, keepdims=True)
        return x / np.sqrt((x * x).mean(1, keepdims=True) + eps)

    inp = sig(h @ np.asarray(w_ig, np.float64).T + np.asarray(b_ig, np.float64)) * m
    gi = inp @ np.asarray(w_ih, np.float64).T
    gh = h @ np.asarray(w_hh, np.float64).T
    if layernorm:
        gi, gh = ln(gi), ln(gh)
    b_ih, b_hh = np.asarray(b_ih, np.float64), np.asarray(b_hh, np.float64)
    r = sig(gi[:, :F] + b_ih[:F] + gh[:, :F] + b_hh[:F])
    z = sig(gi[:, F:2 * F] + b_ih[F:2 * F] + gh[:, F:2 * F] + b_hh[F:2 * F])
    n = np.tanh(gi[:, 2 * F:] + b_ih[2 * F:] + r * (gh[:, 2 * F:] + b_hh[2 * F:]))
    return n + z * (h - n)


def weak_label_propagation(sp_semantic_label, adjacency, sp_semantic_value, sp_pred_semantic, affinity_matrix,
                           iterations_num, class_num=20):
    """Random-walk label propagation, restating modules/datasets/scannetv2_dataset.py:664-735 in numpy
    float64 (the reference itself is numpy float64).  igraph is only the reference's container for
    `adjacency` (get_adjacency(), :679) and the seed labels (vs['semantic_label'], :676).

    Returns (pseudo_label_final float64[S] (-100 = none), pseudo_label_scores float64[S])."""
    lab = np.asarray(sp_semantic_label)
    S = lab.shape[0]
    adj = np.asarray(adjacency, dtype=np.float64) + np.eye(S)                      # :679-682
    aff = np.asarray(affinity_matrix, dtype=np.float64)
    assert adj.shape == aff.shape
    val = np.asarray(sp_semantic_value)
    pred = np.asarray(sp_pred_semantic)
    scores_list, pseudo_list = [], []
    for i in range(class_num):                                                     # :689
        if (lab == i).sum() == 0:
            continue
        sem = np.zeros(adj.shape)
        rows = (pred == i) & (val > 0.7)                                           # :697
        sem[rows] = rows.astype("int")
        for ind, flag in enumerate(lab == i):                                      # :698-700
            if flag:
                sem[ind][ind] = 1
        w = aff * adj * sem                                                        # :702
        d = np.sum(w, axis=1, keepdims=True)                                       # :705
        d[d == 0] += 1
        trans = w / d
        t = trans
        for _ in range(iterations_num):                                            # :710-712
            trans = np.dot(trans, t)
        prob = np.zeros(trans.shape)
        prob[lab == i] = trans[lab == i]                                           # :714-715
        scores_list.append(np.max(prob, axis=0))                                   # :717
        pseudo_list.append(np.argmax(prob, axis=0))                                # :718
    if not scores_list:
        return np.ones(S) * -100, np.zeros(S)
    scores_list = np.array(scores_list)
    pseudo_list = np.array(pseudo_list)
    ind = np.argmax(scores_list, axis=0)                                           # :728
    pseudo = np.take_along_axis(pseudo_list, ind[None], 0)[0]                      # np.choose, :729 (choose caps at 32 rows)
    pscores = np.take_along_axis(scores_list, ind[None], 0)[0]                     # :730
    final = np.ones(S) * -100                                                      # :732
    mask = (pscores != 0) & (lab == -100)                                          # :733
    final[mask] = pseudo[mask]
    return final, pscores


def dense_affinity(edge_u, edge_v, edge_affinity, spnum):
    """train_scannetv2.py:565-570: affinity_matrix[u][v] = aff (float64, later duplicates win)."""
    a = np.zeros((spnum, spnum))
    for u, v, w in zip(edge_u, edge_v, edge_affinity):
        a[u][v] = w
    return a
