"""Deterministic synthetic scenes shaped like the reference's inputs.  There is no network for
ScanNet / S3DIS, so every benchmark and parity test runs on these; seeds are `1000*config + scene_index`.

A scene is what the reference's Dataset.__getitem__ + collate_fn hand to the training loop
(modules/datasets/scannetv2_dataset.py:96-191, 343-474): points (xyz float32 metres, rgb in [-1,1]), integer
voxel coordinates floor(xyz*scale) shifted to >= 0 (:149-153,177), a superpoint id per point (contiguous
0..S-1, :424), a symmetric superpoint graph with 13-d edge features
(data/ScanNetV2/prepare_data_inst_ScanNetV2.py:191-266) and one weak-label seed superpoint per instance
(config `annotation_num: 1`).
"""
import numpy as np
import torch


def _sample_rect(rng, n, origin, eu, ev, normal, jitter):
    """n points on the rectangle origin + a*eu + b*ev, a,b in [0,1], with gaussian jitter along the normal."""
    a, b = rng.random(n), rng.random(n)
    p = origin[None] + a[:, None] * eu[None] + b[:, None] * ev[None]
    p = p + rng.normal(0.0, jitter, n)[:, None] * normal[None]
    return p, a * np.linalg.norm(eu), b * np.linalg.norm(ev)


def make_scene(seed, n_points=150000, room=(8.0, 6.0, 2.6), n_boxes=12, sp_cell=0.25, scale=50, classes=20,
               box_scale=1.0):
    """ScanNet-shaped room: floor + 4 walls + `n_boxes` axis-aligned boxes, sigma = 5 mm jitter."""
    rng = np.random.default_rng(seed)
    X, Y, Z = room
    rects = []  # (origin, eu, ev, normal, instance id, class id)
    rects.append((np.array([0, 0, 0.]), np.array([X, 0, 0.]), np.array([0, Y, 0.]), np.array([0, 0, 1.]), 0, 1))
    walls = [(np.array([0, 0, 0.]), np.array([X, 0, 0.]), np.array([0, 0, Z]), np.array([0, 1., 0])),
             (np.array([0, Y, 0.]), np.array([X, 0, 0.]), np.array([0, 0, Z]), np.array([0, 1., 0])),
             (np.array([0, 0, 0.]), np.array([0, Y, 0.]), np.array([0, 0, Z]), np.array([1., 0, 0])),
             (np.array([X, 0, 0.]), np.array([0, Y, 0.]), np.array([0, 0, Z]), np.array([1., 0, 0]))]
    for w in walls:
        rects.append(w + (1, 0))
    inst = 2
    for _ in range(n_boxes):
        sx, sy, sz = (rng.uniform(0.4, 1.6) * box_scale, rng.uniform(0.4, 1.6) * box_scale,
                      rng.uniform(0.4, 1.2) * box_scale)
        ox, oy = rng.uniform(0.1, X - sx - 0.1), rng.uniform(0.1, Y - sy - 0.1)
        cls = int(rng.integers(2, classes))
        o = np.array([ox, oy, 0.0])
        ex, ey, ez = np.array([sx, 0, 0.]), np.array([0, sy, 0.]), np.array([0, 0, sz])
        faces = [(o + ez, ex, ey, np.array([0, 0, 1.])), (o, ex, ez, np.array([0, 1., 0])),
                 (o + ey, ex, ez, np.array([0, 1., 0])), (o, ey, ez, np.array([1., 0, 0])),
                 (o + ex, ey, ez, np.array([1., 0, 0]))]
        for f in faces:
            rects.append(f + (inst, cls))
        inst += 1
    areas = np.array([np.linalg.norm(np.cross(r[1], r[2])) for r in rects])
    counts = np.floor(areas / areas.sum() * n_points).astype(np.int64)
    counts[0] += n_points - counts.sum()
    xyz, sp_key, inst_id, cls_id = [], [], [], []
    for ri, (r, n) in enumerate(zip(rects, counts)):
        p, a, b = _sample_rect(rng, int(n), r[0], r[1], r[2], r[3], 0.005)
        xyz.append(p)
        sp_key.append(np.stack([np.full(n, ri), np.floor(a / sp_cell).astype(np.int64),
                                np.floor(b / sp_cell).astype(np.int64)], 1))
        inst_id.append(np.full(n, r[4]))
        cls_id.append(np.full(n, r[5]))
    xyz = np.concatenate(xyz).astype(np.float32)
    sp_key = np.concatenate(sp_key)
    inst_id, cls_id = np.concatenate(inst_id), np.concatenate(cls_id)
    perm = rng.permutation(n_points)  # scanners do not deliver points surface by surface
    xyz, sp_key, inst_id, cls_id = xyz[perm], sp_key[perm], inst_id[perm], cls_id[perm]
    rgb = rng.uniform(-1, 1, (n_points, 3)).astype(np.float32)

    uniq, superpoint = np.unique(sp_key, axis=0, return_inverse=True)
    superpoint = superpoint.reshape(-1).astype(np.int64)
    S = uniq.shape[0]
    centers = np.zeros((S, 3))
    np.add.at(centers, superpoint, xyz)
    centers /= np.maximum(np.bincount(superpoint, minlength=S), 1)[:, None]
    sp_inst = np.zeros(S, np.int64)
    sp_cls = np.zeros(S, np.int64)
    sp_inst[superpoint] = inst_id
    sp_cls[superpoint] = cls_id

    # edges: 4-neighbourhood on each surface patch grid + <=5 nearest centres within 0.3 m, symmetric
    key_to_sp = {tuple(k): i for i, k in enumerate(uniq.tolist())}
    edges = set()
    for i, (r, a, b) in enumerate(uniq.tolist()):
        for da, db in ((1, 0), (0, 1)):
            j = key_to_sp.get((r, a + da, b + db))
            if j is not None:
                edges.add((i, j))
                edges.add((j, i))
    from scipy.spatial import cKDTree
    tree = cKDTree(centers)
    dist, nn = tree.query(centers, k=6, distance_upper_bound=0.3)
    for i in range(S):
        for d, j in zip(dist[i, 1:], nn[i, 1:]):
            if np.isfinite(d) and j < S and j != i:
                edges.add((i, int(j)))
                edges.add((int(j), i))
    edges = np.array(sorted(edges), dtype=np.int64).reshape(-1, 2)
    edgefeats = rng.standard_normal((edges.shape[0], 13)).astype(np.float32)

    seed_label = np.full(S, -100, np.int64)
    for i in np.unique(sp_inst):
        cand = np.nonzero(sp_inst == i)[0]
        s = int(cand[rng.integers(0, len(cand))])
        seed_label[s] = sp_cls[s]

    locs = np.floor(xyz * scale).astype(np.int64)
    locs -= locs.min(0)
    return dict(xyz=xyz, rgb=rgb, locs=locs, superpoint=superpoint, edges=edges, edgefeats=edgefeats,
                sp_class=sp_cls, sp_instance=sp_inst, seed_label=seed_label, num_superpoints=S)


def make_room_s3dis(seed, n_points=1000000):
    """S3DIS-shaped room (BASELINE.json configs[3]): 20 m x 15 m x 3 m, 1 M points, 5 cm voxels
    (scale 20), ~10k superpoints (0.4 m patches), 40 boxes."""
    return make_scene(seed, n_points=n_points, room=(20.0, 15.0, 3.0), n_boxes=40, sp_cell=0.4, scale=20,
                      box_scale=1.5)


def make_shell(n_floor=(400, 300), wall=(400, 75)):
    """The reproducible 150 000-voxel micro-shape of SURVEY.md §6: a 400x300 floor plus a 400x75 wall, one
    voxel per cell (spatial_shape [400,300,128])."""
    fx, fy = np.meshgrid(np.arange(n_floor[0]), np.arange(n_floor[1]), indexing="ij")
    floor = np.stack([fx.ravel(), fy.ravel(), np.zeros(fx.size, np.int64)], 1)
    wx, wz = np.meshgrid(np.arange(wall[0]), np.arange(1, wall[1] + 1), indexing="ij")
    w = np.stack([wx.ravel(), np.zeros(wx.size, np.int64), wz.ravel()], 1)
    c = np.concatenate([floor, w]).astype(np.int64)
    return np.concatenate([np.zeros((c.shape[0], 1), np.int64), c], 1), [n_floor[0], n_floor[1], 128]


def training_labels(scene, labelled_fraction=0.3, ignore=-100):
    """Weak labels of one scene in the form the reference's collate_fn hands to the loss (scannetv2_dataset.py:
    362-440): per-superpoint semantic / instance label (-100 = unlabelled; the seed superpoints plus a
    `labelled_fraction` of pseudo-labelled ones, as in the middle of training), the offset from the superpoint centre
    to its instance centre, log(#voxels of the instance) and the instance's size; point labels = the label of the
    point's superpoint."""
    rng = np.random.default_rng(int(scene["num_superpoints"]) * 7919 + 17)
    S = scene["num_superpoints"]
    sp, xyz = scene["superpoint"], scene["xyz"].astype(np.float64)
    cnt = np.maximum(np.bincount(sp, minlength=S), 1)
    centers = np.stack([np.bincount(sp, weights=xyz[:, d], minlength=S) for d in range(3)], 1) / cnt[:, None]
    inst = scene["sp_instance"]
    n_inst = int(inst.max()) + 1
    pinst = inst[sp]
    icnt = np.maximum(np.bincount(pinst, minlength=n_inst), 1)
    icenter = np.stack([np.bincount(pinst, weights=xyz[:, d], minlength=n_inst) for d in range(3)], 1) / icnt[:, None]
    vox = np.unique(np.concatenate([pinst[:, None], scene["locs"]], 1), axis=0)
    ivox = np.maximum(np.bincount(vox[:, 0], minlength=n_inst), 1)
    lo = np.full((n_inst, 3), np.inf)
    hi = np.full((n_inst, 3), -np.inf)
    np.minimum.at(lo, pinst, xyz)
    np.maximum.at(hi, pinst, xyz)
    isize = np.linalg.norm(np.where(np.isfinite(hi - lo), hi - lo, 0.0), axis=1)
    labelled = (scene["seed_label"] != ignore) | (rng.random(S) < labelled_fraction)
    sp_sem = np.where(labelled, scene["sp_class"], ignore).astype(np.int64)
    sp_ins = np.where(labelled, inst, ignore).astype(np.int64)
    return dict(sp_sem=sp_sem, sp_ins=sp_ins, sp_offset=(icenter[inst] - centers).astype(np.float32),
                sp_voxnum=np.log(ivox[inst].astype(np.float32)), sp_size=isize[inst].astype(np.float32),
                pt_sem=sp_sem[sp], pt_ins=sp_ins[sp])


def collate(scenes, full_scale=128, with_labels=False):
    """Batch scenes the way collate_fn does (scannetv2_dataset.py:343-474): batch index in locs[:,0], superpoint
    and edge ids offset per scene, spatial_shape = clip(max+1, full_scale[0], None) (:445), edges for the ECC
    network sorted by target (ecc/GraphConvInfo.py:50-76).  CPU tensors, ready for pin_memory()."""
    locs, feats_rgb, xyz, sp, eu, ev, ef, seed, sp_off = [], [], [], [], [], [], [], [], [0]
    base = 0
    for b, s in enumerate(scenes):
        n = s["locs"].shape[0]
        locs.append(np.concatenate([np.full((n, 1), b, np.int64), s["locs"]], 1))
        feats_rgb.append(s["rgb"])
        xyz.append(s["xyz"])
        sp.append(s["superpoint"] + base)
        eu.append(s["edges"][:, 0] + base)
        ev.append(s["edges"][:, 1] + base)
        ef.append(s["edgefeats"])
        seed.append(s["seed_label"])
        base += s["num_superpoints"]
        sp_off.append(base)
    locs = np.concatenate(locs)
    eu, ev, ef = np.concatenate(eu), np.concatenate(ev), np.concatenate(ef)
    order = np.argsort(ev, kind="stable")
    spatial_shape = np.clip(locs.max(0)[1:] + 1, full_scale, None)
    extra = {}
    if with_labels:
        labs, ibase = [training_labels(s) for s in scenes], 0
        for s, lab in zip(scenes, labs):                      # instance ids are unique across the batch (:386-389)
            for k in ("sp_ins", "pt_ins"):
                lab[k] = np.where(lab[k] >= 0, lab[k] + ibase, lab[k])
            ibase += int(s["sp_instance"].max()) + 1
        cat = lambda k: torch.from_numpy(np.concatenate([lab[k] for lab in labs]))  # noqa: E731
        extra = dict(semantic_labels=cat("pt_sem"), instance_labels=cat("pt_ins"),
                     superpoint_semantic_labels=cat("sp_sem"), superpoint_instance_labels=cat("sp_ins"),
                     superpoint_offset_vector=cat("sp_offset"), superpoint_instance_voxel_num=cat("sp_voxnum"),
                     superpoint_instance_size=cat("sp_size"))
    return dict(extra, 
        locs=torch.from_numpy(locs), locs_float=torch.from_numpy(np.concatenate(xyz)),
        feats=torch.from_numpy(np.concatenate(feats_rgb)), superpoint=torch.from_numpy(np.concatenate(sp)),
        edge_u_list=torch.from_numpy(eu), edge_v_list=torch.from_numpy(ev),
        ecc_edge_index=torch.from_numpy(np.stack([eu[order], ev[order]])), ecc_edgefeats=torch.from_numpy(ef[order]),
        seed_label=torch.from_numpy(np.concatenate(seed)), sp_batch_offsets=torch.tensor(sp_off, dtype=torch.int32),
        spatial_shape=[int(x) for x in spatial_shape], batch_size=len(scenes), num_superpoints=base)
