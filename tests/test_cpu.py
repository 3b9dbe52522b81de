"""CPU suite (`pytest -m "not gpu"`): pins the ORACLE against the reference's golden vectors and against the
reference's own compiled CPU kernels, checks the C-ABI library's surface, the CUDA-free host routine
(voxelization_idx for DataLoader workers), the drop-in module API, and the multi-process host logic (gloo, world 2).
No CUDA compute call happens here."""
import ast
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-wsis_b200")


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)


def _scene_batch(z):
    from wsis_b200 import synthetic
    nscene, kw = ast.literal_eval(str(z["scene_kw"]))
    return synthetic.collate([synthetic.make_scene(9000 + i, **kw) for i in range(nscene)])


def _sorted_pairs(pairs, num, in_coords, out_coords):
    """Canonical form of a rulebook: for every offset the sorted set of (in coord, out coord) rows."""
    out = []
    for k in range(pairs.shape[0]):
        n = int(num[k])
        rows = np.concatenate([in_coords[pairs[k, 0, :n]], out_coords[pairs[k, 1, :n]]], 1)
        out.append(rows[np.lexsort(rows.T[::-1])] if n else rows)
    return out


# ---------------------------------------------------------------------------------------------------------
# oracle vs golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["scene_b1", "scene_b2"])
def test_oracle_rulebooks_match_reference_golden(orc, golden_dir, name):
    z = _golden(golden_dir, name)
    batch = _scene_batch(z)
    bs, shape = batch["batch_size"], batch["spatial_shape"]
    coords = z["voxel_locs"].astype(np.int32)
    assert np.array_equal(z["rb_subm1_outids"], coords)
    pairs, num = orc.rulebook_subm(coords, bs, shape, 3, 1)
    assert np.array_equal(num, z["rb_subm1_num"])
    assert np.array_equal(pairs, z["rb_subm1_pairs"])            # CPU reference order is deterministic: bit-exact
    oc, cp, cn, oshape = orc.rulebook_conv(coords, bs, shape, 2, 2, 0, 1)
    assert np.array_equal(oc, z["rb_spconv1_outids"])
    assert np.array_equal(cn, z["rb_spconv1_num"])
    assert np.array_equal(cp, z["rb_spconv1_pairs"])
    _, num2 = orc.rulebook_subm(oc, bs, oshape, 3, 1)
    assert np.array_equal(num2, z["rb_subm2_num"])
    oc2, _, cn2, _ = orc.rulebook_conv(oc, bs, oshape, 2, 2, 0, 1)
    assert np.array_equal(oc2, z["rb_spconv2_outids"]) and np.array_equal(cn2, z["rb_spconv2_num"])


@pytest.mark.parametrize("name", ["scene_b1", "scene_b2"])
def test_oracle_voxelization_matches_golden(orc, golden_dir, name):
    z = _golden(golden_dir, name)
    batch = _scene_batch(z)
    locs, p2v, v2p = orc.voxelization_idx(batch["locs"].numpy(), batch["batch_size"], 4)
    assert np.array_equal(locs, z["voxel_locs"]) and np.array_equal(p2v, z["p2v"]) and np.array_equal(v2p, z["v2p"])
    feats = torch.cat((batch["feats"], batch["locs_float"]), 1).numpy()
    assert np.array_equal(orc.voxelization(feats, v2p, 4), z["voxel_feats"])
    centers = orc.scatter(batch["locs_float"].numpy(), batch["superpoint"].numpy(), "mean")
    np.testing.assert_allclose(centers, z["centers"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's own CPU kernels compiled from /root/reference (oracle/_ref), where present
# ---------------------------------------------------------------------------------------------------------
def _ref():
    from oracle import ref_spconv
    if not ref_spconv.available():
        pytest.skip("oracle/_ref/libspconv_ref.so not built (needs /root/reference)")
    return ref_spconv


def _coords(rng, shape, n, bs):
    cells = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, 3)
    out = []
    for b in range(bs):
        sel = cells[rng.permutation(len(cells))[:n]]
        out.append(np.concatenate([np.full((n, 1), b), sel], 1))
    return np.concatenate(out).astype(np.int32)


@pytest.mark.parametrize("bs,k,s,p,d,subm", [(1, 3, 1, 1, 1, True), (2, 3, 1, 1, 1, True), (2, 3, 1, 2, 2, True),
                                             (1, 2, 2, 0, 1, False), (2, 3, 2, 1, 1, False), (2, 3, 1, 0, 2, False),
                                             (1, 3, 3, 2, 1, False), (2, 2, 1, 1, 1, False)])
def test_oracle_rulebook_equals_compiled_reference(orc, bs, k, s, p, d, subm):
    ref = _ref()
    rng = np.random.default_rng(100 * k + 10 * s + p + d + bs)
    shape = [19, 18, 17]                                       # the reference's own test grid (test_conv.py:329)
    coords = _coords(rng, shape, 700, bs)
    outids, rp, rn = ref.get_indice_pairs(torch.from_numpy(coords), bs, shape, k, s, p, d, 0, subm)
    if subm:
        pairs, num = orc.rulebook_subm(coords, bs, shape, k, d)
        oc = coords
    else:
        oc, pairs, num, _ = orc.rulebook_conv(coords, bs, shape, k, s, p, d)
    assert np.array_equal(oc, outids.numpy())
    assert np.array_equal(num, rn.numpy())
    assert np.array_equal(pairs, rp.numpy())


@pytest.mark.parametrize("kind", ["subm", "conv", "inverse"])
def test_oracle_conv_fwd_bwd_equals_compiled_reference(orc, kind):
    ref = _ref()
    rng = np.random.default_rng(11)
    shape, bs, cin, cout = [19, 18, 17], 2, 24, 40
    coords = _coords(rng, shape, 600, bs)
    if kind == "subm":
        pairs, num = orc.rulebook_subm(coords, bs, shape, 3, 1)
        n_in = n_out = coords.shape[0]
        w = rng.uniform(-.3, .3, (3, 3, 3, cin, cout)).astype(np.float32)
    else:
        oc, pairs, num, _ = orc.rulebook_conv(coords, bs, shape, 2, 2, 0, 1)
        n_in, n_out = coords.shape[0], oc.shape[0]
        w = rng.uniform(-.3, .3, (2, 2, 2, cin, cout)).astype(np.float32)
    inverse = kind == "inverse"
    if inverse:
        n_in, n_out = n_out, n_in
    x = rng.uniform(-1, 1, (n_in, cin)).astype(np.float32)
    g = rng.uniform(-1, 1, (n_out, cout)).astype(np.float32)
    y = orc.indice_conv(x, w, pairs, num, n_out, inverse)
    t = torch.from_numpy
    yr = ref.indice_conv(t(x), t(w), t(pairs), t(num), n_out, inverse, kind == "subm").numpy()
    np.testing.assert_allclose(y, yr, rtol=1e-5, atol=1e-5)
    din, dw = orc.indice_conv_backward(x, w, g, pairs, num, inverse)
    dinr, dwr = ref.indice_conv_backward(t(x), t(w), t(g), t(pairs), t(num), inverse, kind == "subm")
    np.testing.assert_allclose(din, dinr.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dw, dwr.numpy().reshape(dw.shape), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------------
# the reference's own known-answer test: sparse conv == dense conv on the active sites
# (modules/lib/spconv/test/test_conv.py:325-382: seed 484, 19x18x17 grid, atol 1e-4)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,s,p,d", [(2, 1, 0, 1), (2, 2, 0, 1), (3, 1, 1, 1), (3, 2, 1, 1), (3, 3, 2, 1), (3, 1, 2, 2),
                                     (3, 1, 0, 3)])
def test_oracle_sparse_conv_equals_dense_conv3d(orc, k, s, p, d):
    np.random.seed(484)
    rng = np.random.default_rng(484)
    shape, bs, cin, cout = [19, 18, 17], 2, 16, 12
    coords = _coords(rng, shape, 1000, bs)
    x = rng.uniform(-1, 1, (coords.shape[0], cin)).astype(np.float32)
    w = rng.uniform(-1, 1, (k, k, k, cin, cout)).astype(np.float32)
    oc, pairs, num, oshape = orc.rulebook_conv(coords, bs, shape, k, s, p, d)
    y = orc.indice_conv(x, w, pairs, num, oc.shape[0])
    dense = torch.zeros(bs, cin, *shape)
    dense[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]] = torch.from_numpy(x)
    wd = torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous()
    yd = torch.nn.functional.conv3d(dense, wd, stride=s, padding=p, dilation=d)
    assert list(yd.shape[2:]) == oshape
    got = torch.zeros_like(yd)
    got[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]] = torch.from_numpy(y)
    # every cell the dense conv makes non-zero must be an active output, and the values must agree
    np.testing.assert_allclose(got.numpy(), yd.numpy(), atol=1e-4)


def test_oracle_subm_equals_dense_conv_sampled_at_active_sites(orc):
    rng = np.random.default_rng(5)
    shape, bs, cin, cout = [19, 18, 17], 2, 8, 8
    coords = _coords(rng, shape, 900, bs)
    x = rng.uniform(-1, 1, (coords.shape[0], cin)).astype(np.float32)
    w = rng.uniform(-1, 1, (3, 3, 3, cin, cout)).astype(np.float32)
    pairs, num = orc.rulebook_subm(coords, bs, shape, 3, 1)
    y = orc.indice_conv(x, w, pairs, num, coords.shape[0])
    dense = torch.zeros(bs, cin, *shape)
    dense[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]] = torch.from_numpy(x)
    yd = torch.nn.functional.conv3d(dense, torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous(), padding=1)
    np.testing.assert_allclose(y, yd[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]].numpy(), atol=1e-4)


def test_oracle_conv_then_inverse_restores_indices(orc):
    """k2 s2 conv + inverse conv is the UNet's down/up couple (test_conv.py:443-497 checks it against
    SparseConvNet, which is not installed): the inverse rulebook is the same pair list read backwards, so the
    up-sampled tensor lives on exactly the original voxels and each receives exactly one contribution."""
    rng = np.random.default_rng(9)
    shape, bs = [19, 18, 17], 2                               # odd extents: the last plane has no k2s2 output
    coords = _coords(rng, shape, 800, bs)
    oc, pairs, num, oshape = orc.rulebook_conv(coords, bs, shape, 2, 2, 0, 1)
    assert oshape == [9, 9, 8]
    covered = np.zeros(coords.shape[0], np.int64)
    for k in range(8):
        np.add.at(covered, pairs[k, 0, :num[k]], 1)
    inside = np.all(coords[:, 1:] < 2 * np.asarray(oshape), axis=1)
    assert np.array_equal(covered, inside.astype(np.int64))     # dropped exactly where the odd extent cuts
    ones = np.ones((oc.shape[0], 4), np.float32)
    w = np.tile(np.eye(4, dtype=np.float32), (2, 2, 2, 1, 1))
    up = orc.indice_conv(ones, w, pairs, num, coords.shape[0], inverse=True)
    assert np.array_equal(up[:, 0], inside.astype(np.float32))


def test_oracle_edge_cases(orc):
    shape = [8, 8, 8]
    empty = np.zeros((0, 4), np.int32)
    pairs, num = orc.rulebook_subm(empty, 1, shape, 3, 1)
    assert pairs.shape == (27, 2, 0) and num.sum() == 0
    oc, pairs, num, _ = orc.rulebook_conv(empty, 1, shape, 2, 2, 0, 1)
    assert oc.shape == (0, 4) and num.sum() == 0
    one = np.array([[0, 7, 7, 7]], np.int32)                  # corner voxel: only the centre offset pairs up
    pairs, num = orc.rulebook_subm(one, 1, shape, 3, 1)
    assert num.tolist() == [0] * 13 + [1] + [0] * 13 and pairs[13, :, 0].tolist() == [0, 0]
    full = _coords(np.random.default_rng(0), [4, 4, 4], 64, 1)  # fully occupied grid: interior voxels see 27
    pairs, num = orc.rulebook_subm(full, 1, [4, 4, 4], 3, 1)
    assert num[13] == 64 and num.sum() == sum((4 - abs(a)) * (4 - abs(b)) * (4 - abs(c))
                                               for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1))


def test_oracle_random_walk_properties(orc):
    """scannetv2_dataset.py:664-735 restated in numpy float64: seeds never get a pseudo label, labels are seed
    superpoint ids of a seeded class, zero iterations = one hop, scores are probabilities."""
    rng = np.random.default_rng(3)
    S = 60
    adj = np.zeros((S, S))
    for i in range(S - 1):
        adj[i, i + 1] = adj[i + 1, i] = 1
    aff = rng.uniform(0.1, 1.0, (S, S))
    lab = np.full(S, -100)
    lab[[5, 40]] = [2, 7]
    pred = np.where(np.arange(S) < 30, 2, 7)
    conf = np.full(S, 0.9)
    for it in (0, 1, 3):
        final, score = orc.weak_label_propagation(lab, adj, conf, pred, aff, it)
        assert np.all(final[[5, 40]] == -100)
        got = set(np.unique(final[final != -100]).astype(int).tolist())
        assert got <= {5, 40}
        assert np.all((score >= 0) & (score <= 1 + 1e-12))
        reach = (final != -100).sum()
        assert reach >= (2 if it == 0 else 4)
    final0, _ = orc.weak_label_propagation(lab, adj, conf, pred, aff, 0)
    assert set(np.nonzero(final0 != -100)[0].tolist()) == {4, 6, 39, 41}   # one hop from each seed


def test_oracle_random_walk_matches_reference_golden(orc, golden_dir):
    """The oracle's restatement against the reference's OWN `weak_label_propagation`
    (modules/datasets/scannetv2_dataset.py:664-735, cut out with ast and executed by tests/golden/make_golden_rw.py)
    on a 150k-point scene's superpoint graph (S = 3138, 16 288 directed edges), iterations_num 0, 1 and 3: pseudo labels
    identical, scores bit-equal (both are numpy float64 doing the same operations in the same order)."""
    g = np.load(os.path.join(golden_dir, "rw_scene0.npz"))
    S, e = int(g["S"]), g["edges"].astype(np.int64)
    adj = np.zeros((S, S))
    adj[e[:, 0], e[:, 1]] = 1
    A = orc.dense_affinity(e[:, 0], e[:, 1], g["aff"], S)
    for it in (0, 1, 3):
        final, score = orc.weak_label_propagation(g["seed_label"].astype(np.int64), adj, g["conf"], g["pred"].astype(np.int64), A, it)
        assert np.array_equal(final.astype(np.int32), g["pseudo_it%d" % it])
        assert np.array_equal(score, g["score_it%d" % it])
        assert (final != -100).sum() > (10, 20, 0, 40)[it]


# ---------------------------------------------------------------------------------------------------------
# the C-ABI library: loads, exports every declared symbol, pure helpers answer without a GPU
# ---------------------------------------------------------------------------------------------------------
def test_cabi_library_exports_every_declared_symbol():
    from wsis_b200._lib import HEADER_PATH, LIB_PATH, lib, parse_header
    assert os.path.exists(LIB_PATH), "build it first: python __graft_entry__.py build"
    protos = parse_header()
    declared = set(re.findall(r"\b(wsis_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", open(HEADER_PATH).read(), flags=re.S)))
    assert declared == set(protos), "header parser missed %s" % (declared ^ set(protos))
    assert len(protos) >= 35
    dll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        getattr(dll, name)                                     # AttributeError = symbol missing from the .so
    L = lib()
    assert L.value("wsis_version") >= 100
    assert L.value("wsis_hash_slots", 1000) == 2048 and L.value("wsis_hash_slots", 1) == 1024
    assert L.value("wsis_tile_pad", 0) == 0 and L.value("wsis_tile_pad", 129) == 256
    assert L.value("wsis_conv_umma_supported", 32, 32) == 1 and L.value("wsis_conv_umma_supported", 6, 32) == 1
    assert L.value("wsis_conv_umma_supported", 32, 24) == 0 and L.value("wsis_conv_pack_bytes", 27, 6, 32, 1) == 27 * 2 * 32 * 64
    assert L.value("wsis_conv_pack_bytes", 27, 64, 32, 3) == 27 * 2 * 2 * 32 * 64
    assert L.value("wsis_scan_ws_bytes", 1 << 20) > 0 and L.value("wsis_sort_ws_bytes", 1 << 20) > 0
    nm = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln and ln.split()[-1].startswith("wsis_")}
    assert set(protos) <= exported


def test_no_cpu_fallback_for_device_ops():
    from wsis_b200 import ops as W
    x = torch.zeros(4, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        W.sparse_conv(x, torch.zeros(27, 32, 32), torch.zeros(4, 27, dtype=torch.int32), 4, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        W.rulebook_subm(torch.zeros(4, 4, dtype=torch.int32), [8, 8, 8])
    with pytest.raises(RuntimeError, match="CUDA"):
        W.scatter(x, torch.zeros(4, dtype=torch.int64), reduce="mean", dim_size=2)


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|wsis_oracle|libspconv_ref", txt, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, "product files reference the oracle: %s" % bad


# ---------------------------------------------------------------------------------------------------------
# CUDA-free host routine: voxelization_idx on CPU tensors (DataLoader workers, scannetv2_dataset.py:449)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,span,bs", [(0, 4, 1), (1, 4, 1), (5000, 20, 2), (60000, 64, 4), (3000, 3, 1)])
def test_host_voxelization_idx_matches_oracle(orc, n, span, bs):
    import pointgroup_ops
    rng = np.random.default_rng(n + span)
    coords = np.concatenate([rng.integers(0, bs, (n, 1)), rng.integers(0, span, (n, 3))], 1).astype(np.int64)
    locs, p2v, v2p = pointgroup_ops.voxelization_idx(torch.from_numpy(coords), bs, 4)
    assert locs.dtype == torch.int64 and p2v.dtype == torch.int32 and v2p.dtype == torch.int32
    if n == 0:
        assert locs.shape == (0, 4) and p2v.shape == (0,)
        return
    el, ep, ev = orc.voxelization_idx(coords, bs, 4)
    assert np.array_equal(locs.numpy(), el) and np.array_equal(p2v.numpy(), ep) and np.array_equal(v2p.numpy(), ev)
    # contract properties: p2v indexes voxels of the point's own coords; counts add up; lists are point ids
    assert np.array_equal(el[ep], coords)
    assert ev[:, 0].sum() == n and ev[:, 0].max() == ev.shape[1] - 1


def test_host_voxelization_idx_golden(golden_dir):
    import pointgroup_ops
    for name in ("scene_b1", "scene_b2"):
        z = _golden(golden_dir, name)
        batch = _scene_batch(z)
        locs, p2v, v2p = pointgroup_ops.voxelization_idx(batch["locs"], batch["batch_size"], 4)
        assert np.array_equal(locs.numpy(), z["voxel_locs"]) and np.array_equal(p2v.numpy(), z["p2v"])
        assert np.array_equal(v2p.numpy(), z["v2p"])


# ---------------------------------------------------------------------------------------------------------
# drop-in module API: names, parameter layout, construction order (fixed-seed weights == the reference's)
# ---------------------------------------------------------------------------------------------------------
def test_spconv_api_surface():
    import spconv
    from spconv.modules import SparseModule
    for name in ("SparseConvTensor", "SparseSequential", "SubMConv3d", "SparseConv3d", "SparseInverseConv3d",
                 "SparseModule", "ops", "functional"):
        assert hasattr(spconv, name), name
    for fn in ("get_indice_pairs", "indice_conv", "indice_conv_backward", "get_conv_output_size",
               "get_deconv_output_size"):
        assert hasattr(spconv.ops, fn)
    for fn in ("indice_conv", "indice_inverse_conv", "indice_subm_conv"):
        assert hasattr(spconv.functional, fn)
    conv = spconv.SubMConv3d(6, 32, kernel_size=3, padding=1, bias=False, indice_key="subm1")
    assert tuple(conv.weight.shape) == (3, 3, 3, 6, 32) and conv.bias is None and isinstance(conv, SparseModule)
    down = spconv.SparseConv3d(32, 64, kernel_size=2, stride=2, bias=False, indice_key="spconv1")
    assert tuple(down.weight.shape) == (2, 2, 2, 32, 64) and not down.subm and not down.inverse
    up = spconv.SparseInverseConv3d(64, 32, kernel_size=2, bias=False, indice_key="spconv1")
    assert up.inverse and tuple(up.weight.shape) == (2, 2, 2, 64, 32)
    with pytest.raises(AssertionError):
        spconv.SparseConv3d(4, 4, 3, stride=2, dilation=2)       # conv.py:80-81
    t = spconv.SparseConvTensor(torch.zeros(3, 2), torch.tensor([[0, 0, 0, 0], [0, 1, 1, 1], [0, 2, 0, 1]],
                                                                  dtype=torch.int32), [3, 2, 2], 1)
    assert t.spatial_size == 12 and abs(t.sparity - 0.25) < 1e-12 and t.find_indice_pair("x") is None
    assert tuple(t.dense().shape) == (1, 2, 3, 2, 2)
    seq = spconv.SparseSequential(torch.nn.Linear(2, 4), torch.nn.ReLU())
    assert len(seq) == 2 and isinstance(seq[1], torch.nn.ReLU)
    out = seq(t)                                                    # dense modules act on .features in place
    assert out is t and tuple(t.features.shape) == (3, 4)
    assert spconv.ops.get_conv_output_size([19, 18, 17], [2] * 3, [2] * 3, [0] * 3, [1] * 3) == [9, 9, 8]


def test_reference_model_files_run_on_the_dropin():
    """The boundary claim of DESIGN.md 1 / INTEGRATION.md 1: the reference's own model files
    (modules/model/backbone_3D_WSIS.py, sparse_unet3d.py, graphnet.py, spg_modules.py), imported unmodified, build
    their Network on THIS repository's `spconv` package and end up with exactly the mirror's state_dict (same keys
    in the same order, bit-equal tensors under the same seed).  Needs the reference tree (this container only)."""
    import json
    ref = os.environ.get("WSIS_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "modules/model/backbone_3D_WSIS.py")):
        pytest.skip("reference tree not present on this box")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_dropin_check.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["same_keys"] and res["n_keys"] > 300 and res["different_tensors"] == []
    assert res["uses_dropin"] and res["ublock_is_reference"] and res["sparse_conv_modules"] == 49


def test_network_parameters_match_reference_golden(golden_dir):
    """Same module/parameter names and shapes as the reference Network (a released state_dict loads strict), and the
    same RNG consumption order: weights built under torch.manual_seed(123) equal the reference's, parameter by
    parameter (checksums recorded by make_golden.py from the reference's own classes)."""
    from wsis_b200 import pipeline
    net = pipeline.build_network(seed=123, device="cpu")
    keys = np.load(os.path.join(golden_dir, "state_dict_keys.npz"))
    sd = net.state_dict()
    assert list(sd.keys()) == keys["keys"].tolist()
    assert [repr(tuple(v.shape)) for v in sd.values()] == keys["shapes"].tolist()
    z = _golden(golden_dir, "scene_b1")
    names = [n for n, _ in net.named_parameters()]
    assert names == z["param_names"].tolist()
    # BatchNorm parameters were overwritten after construction by make_golden.py; all others are init-time values
    mine = np.array([float(p.detach().double().abs().sum()) for p in net.parameters()])
    conv_like = np.array([p.dim() >= 2 for p in net.parameters()])   # conv / linear weights (1-D = BN, biases)
    np.testing.assert_allclose(mine[conv_like], z["param_checksum"][conv_like], rtol=1e-6)
    assert conv_like.sum() >= 70


# ---------------------------------------------------------------------------------------------------------
# multi-process host logic: gloo, world_size 2 (scene sharding, result gather, one-bucket gradient all-reduce)
# ---------------------------------------------------------------------------------------------------------
_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "3d-wsis_b200"))
import torch, torch.distributed as dist
from wsis_b200 import dist as wd
rank, world = wd.init(backend="gloo")
assert world == 2 and wd.world() == (rank, 2)
mine = wd.shard_scenes(7)
assert mine == list(range(rank, 7, 2))
res = wd.gather_scene_results({i: ("scene", i, rank) for i in mine}, 7)
assert [r[1] for r in res] == list(range(7)) and [r[2] for r in res] == [0, 1, 0, 1, 0, 1, 0]
torch.manual_seed(0)                                   # identical replicas
model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
for p in model[3].parameters():                        # an unused head (losses_3D_WSIS.py:68)
    pass
bucket = wd.GradBucket(model.parameters())
x = torch.full((4, 5), float(rank + 1))
bucket.zero()
model[2](model[1](model[0](x))).sum().backward()       # model[3] receives no gradient
local = bucket.flat.clone()
bucket.allreduce()
both = [torch.zeros_like(local) for _ in range(2)]
dist.all_gather(both, local)
assert torch.allclose(bucket.flat, (both[0] + both[1]) / 2, atol=1e-6)
assert model[0].weight.grad.data_ptr() == bucket.flat.data_ptr()
assert float(model[3].weight.grad.abs().sum()) == 0.0
opt = torch.optim.SGD(model.parameters(), lr=0.1)
opt.step()
w = [torch.zeros_like(model[0].weight) for _ in range(2)]
dist.all_gather(w, model[0].weight.data)
assert torch.equal(w[0], w[1]), "replicas diverged after the averaged step"
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank " + str(rank) + " ok\n")
sys.stdout.flush()
'''


def test_gloo_world2_sharding_and_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2 and "0" in r.stdout and "1" in r.stdout, r.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the agreed keys (bounded sample, CPU only)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--points", "20000", "--scenes", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "scenes/s" and line["value"] > 0
    # the line reports the steps it really ran, each over the whole (small) batch: value = scenes / time
    assert line["steps"] == 2 and line["warmup"] == 1 and line["config"]["scenes_per_step"] == 2
    assert abs(line["value"] - 2 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


# ---------------------------------------------------------------------------------------------------------
# instance clustering: identical masks against the reference's own function
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cluster_scene0", "cluster_scene1"])
def test_clustering_matches_reference_golden(name):
    """tests/golden/make_golden_cluster.py ran the reference's clustering_in_graph (test_scannetv2.py:281-455) on
    these inputs; the aggregate-based implementation must give the same instances: identical masks and labels, and
    the same confidences."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_cluster as mg
    from wsis_b200 import cluster
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    case, nbrs = mg.make_case(int(gold["seed"]), int(gold["n_points"]))
    for k in ("sem", "off", "occ", "size"):                      # the stored inputs are the ones the reference saw
        assert np.array_equal(case[k], gold[k]), k
    conf, label, masks = cluster.clustering_in_graph(case["xyz"], case["superpoint"], nbrs, case["sem"], case["off"],
                                                     case["occ"], case["size"])
    assert masks.shape[1] == len(case["xyz"]) and masks.dtype == int
    assert np.array_equal(np.packbits(masks.astype(bool), axis=1), gold["masks"])
    assert np.array_equal(label, gold["label_id"])
    assert np.allclose(conf, gold["conf"], rtol=0, atol=0)
    assert masks.sum(0).max() <= 1                               # instances are disjoint


def test_clustering_fragments_and_edge_cases():
    """Two same-class boxes far apart stay two instances; a tiny same-class group whose voxel count is below 0.3 x
    the predicted occupancy is a fragment and is absorbed by the nearest primary of its class when it is within the
    primary's radius; classes outside the instance label set never form instances."""
    from wsis_b200 import cluster
    rng = np.random.default_rng(5)
    pts, sp = [], []

    def blob(center, n, sid):
        pts.append(center + rng.uniform(-0.2, 0.2, (n, 3)))
        sp.append(np.full(n, sid))

    blob(np.array([0.0, 0, 0]), 400, 0)
    blob(np.array([0.3, 0, 0]), 400, 1)        # same object as 0
    blob(np.array([5.0, 0, 0]), 400, 2)        # second object, same class
    blob(np.array([0.6, 0, 0]), 3, 3)          # fragment next to object A (3 points)
    blob(np.array([9.0, 9, 0]), 300, 4)        # floor class: not an instance class
    xyz = np.concatenate(pts).astype(np.float32)
    sp = np.concatenate(sp)
    nbrs = cluster.neighbors_from_edges(np.array([[0, 1], [1, 3], [2, 4]]), 5)
    sem = np.array([4, 4, 4, 4, 1])            # semantic_ind2label[1] = 2 -> not in the instance label set
    off = np.zeros((5, 3), np.float32)
    off[1] = [-0.3, 0, 0]                      # superpoint 1 votes for the centre of superpoint 0
    off[3] = [-0.1, 0, 0]                      # the fragment's centre stays 0.5 m from A's: not BFS-merged (0.25*1.2)
    occ = np.log(np.array([400.0, 400, 400, 400, 300])).astype(np.float32)
    size = np.full(5, 1.2, np.float32)
    conf, label, masks = cluster.clustering_in_graph(xyz, sp, nbrs, sem, off, occ, size)
    assert len(conf) == 2 and list(label) == [5, 5]              # semantic_ind2label[4]
    a, b = masks[0].astype(bool), masks[1].astype(bool)
    assert set(np.unique(sp[a])) == {0, 1, 3} and set(np.unique(sp[b])) == {2}   # the fragment was absorbed by A
    assert not masks[:, sp == 4].any()
    # without any primary instance the reference would index an empty list; here the result is simply empty
    conf, label, masks = cluster.clustering_in_graph(xyz, sp, nbrs, np.array([1, 1, 1, 1, 1]), off, occ, size)
    assert len(conf) == 0 and masks.shape == (0, len(xyz))


def test_conv_umma_launch_plan_exists_for_every_layer_shape():
    """Host-only: every (Cin, Cout, K, precision) the tensor-core conv accepts has a pipeline that fits the 227 KB of
    shared memory and the 512 TMEM columns, and the invariants the kernel's barrier protocol relies on hold (stages are
    a power of two and a multiple of the builder groups and of the issuers: every stage has one owner of each kind)."""
    from wsis_b200._lib import lib
    L = lib()
    shapes = [(6, 32), (32, 32), (64, 32), (32, 64), (64, 64), (128, 64), (96, 96), (192, 96), (128, 128), (256, 128),
              (160, 160), (320, 160), (160, 128), (1, 16), (1024, 256), (33, 240)]
    for cin, cout in shapes:
        for K in (1, 8, 27, 32):
            for prec in (1, 3):
                plan = (ctypes.c_int32 * 11)()
                L.call("wsis_conv_umma_plan", K, cin, cout, prec, plan)
                smem, na, nrc, nrec, nbg, nacc, nmma, nbuf, resident, nwp, us = list(plan)
                assert 0 < smem <= 227 * 1024, (cin, cout, K, prec, smem)
                assert na in (2, 4, 8) and nrc >= 1 and nrec >= 2 and nbuf in (1, 2) and nwp in (1, 2) and 1 <= us <= 4
                assert na % nbg == 0 and na % nmma == 0 and nacc == nmma and nmma in (1, 2, 4)
                assert nbuf * nacc * cout + 32 * na * us <= 512       # accumulators + operand slots fit TMEM
    # the level-1 layers keep their whole packed weight in shared memory and run four issuers
    plan = (ctypes.c_int32 * 11)()
    L.call("wsis_conv_umma_plan", 27, 32, 32, 3, plan)
    assert plan[1] == 8 and plan[6] == 4 and plan[8] == 1 and plan[10] == 1
    L.call("wsis_conv_umma_plan", 27, 64, 64, 3, plan)
    assert plan[1] == 8 and plan[6] == 2 and plan[7] == 2 and plan[8] == 0 and plan[10] == 1
    with pytest.raises(RuntimeError):
        L.call("wsis_conv_umma_plan", 27, 32, 24, 3, plan)      # Cout not a multiple of 16


def test_oracle_ecc_gru_step_matches_module(orc):
    """The numpy restatement of one ECC-GRU step (oracle.ecc_gru_step) against the torch formulation of
    spg_modules.py:97-121 / 226-253 in wsis_b200.model (which the golden network outputs pin to the reference)."""
    from wsis_b200 import model as M
    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    S, E = 60, 300
    for layernorm in (True, False):
        cell = M.GRUCellEx(32, 32, bias=True, layernorm=layernorm, ingate=True).double()
        nn_ = M.NNConv(32, 32)
        src, tgt = rng.integers(0, S, E), rng.integers(0, S - 5, E)
        h = torch.from_numpy(rng.standard_normal((S, 32)))
        w = torch.from_numpy(rng.standard_normal((E, 1024)) * 0.2)
        ei = torch.from_numpy(np.stack([src, tgt]))
        with torch.no_grad():
            ref = cell(nn_(h, ei, w), h).numpy()
        out = orc.ecc_gru_step(h.numpy(), w.numpy(), src, tgt, cell._modules["ig"].weight.detach().numpy(),
                               cell._modules["ig"].bias.detach().numpy(), cell.weight_ih.detach().numpy(),
                               cell.weight_hh.detach().numpy(), cell.bias_ih.detach().numpy(),
                               cell.bias_hh.detach().numpy(), layernorm=layernorm)
        assert np.abs(out - ref).max() < 1e-10


def test_conv_umma_barrier_protocol_model():
    """tools/protocol_model.py transcribes the control flow of every role of conv_umma_kernel (ring indices, parities,
    arrival counts) and checks under random schedules that no parity wait passes before its logical generation has
    completed, that every consumer finds the buffer contents it expects, and that nothing deadlocks -- for every
    launch plan the host planner can produce.  A ring with fewer stages than issuers (an issuer whose next stage is two
    generations ahead on the same buffer) must be caught."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import protocol_model as pm
    from wsis_b200._lib import lib
    L = lib()
    plans = set()
    for cin, cout in [(6, 32), (32, 32), (64, 64), (96, 96), (128, 128), (160, 160), (320, 160), (1024, 256), (32, 16)]:
        for K in (8, 27, 32):
            for prec in (1, 3):
                plan = (ctypes.c_int32 * 11)()
                L.call("wsis_conv_umma_plan", K, cin, cout, prec, plan)
                _, na, nrc, nrec, nbg, _, nmma, nbuf, resident, nwp, us = list(plan)
                plans.add((na, nrc, nrec, nbg, nmma, nbuf, resident, nwp, us))
    assert len(plans) >= 3
    rng = random.Random(5)
    for na, nrc, nrec, nbg, nmma, nbuf, resident, nwp, us in sorted(plans):
        kw = dict(nbuf=nbuf, resident=bool(resident), nwp=nwp, us=us)
        for seed in range(10):
            tiles = pm.random_tiles(rng, rng.randint(1, 7))
            assert pm.Cta(tiles, na, nrc, nrec, nbg, nmma, **kw).run(seed)
        assert pm.Cta([(1, 1)], na, nrc, nrec, nbg, nmma, **kw).run(0)              # a single one-unit tile
        assert pm.Cta([(27, 5)] * 3, na, nrc, nrec, nbg, nmma, **kw).run(1)         # the widest layer
    with pytest.raises(pm.ProtocolError):                                    # four issuers on two stages
        for seed in range(50):
            pm.Cta(pm.random_tiles(rng, 6), 2, 2, 2, 2, 4).run(seed)


def test_label_propagation_matches_reference_golden():
    """tests/golden/make_golden_labels.py ran the reference's extend_label_to_neighbor / propagate_label_to_neighbor /
    propagate_label_to_whole_scene (scannetv2_dataset.py:779-958) on this scene; the array-based versions must give the
    same labels, the same offsets and the same per-edge is1ins."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_labels as mg
    from wsis_b200 import labels as LB
    gold = np.load(os.path.join(ROOT, "tests", "golden", "labels_scene0.npz"))
    case = mg.make_case(int(gold["seed"]), int(gold["n_points"]))
    geo = LB.SuperpointGeometry(case["xyz"], case["superpoint"])
    start = LB.SuperpointLabels(case["sem"], case["ins"], case["off"])

    def check(res, key, edges=True):
        assert np.array_equal(res.semantic, gold[key + "_sem"]) and np.array_equal(res.instance, gold[key + "_ins"])
        assert np.array_equal(res.offset, gold[key + "_off"])
        if edges:
            assert np.array_equal(LB.edge_same_instance(case["edges"], res), gold[key + "_is1ins"])

    ext = LB.extend_label_to_neighbor(start, geo, case["nbrs"], case["conf"], case["pred"])
    check(ext, "ext")
    assert (ext.semantic != -100).sum() > (start.semantic != -100).sum()
    prop = LB.propagate_label_to_neighbor(ext, geo, case["nbrs"], case["conf"], case["pred"])
    check(prop, "prop")
    whole = LB.propagate_label_to_whole_scene(start, geo, case["pred"], case["pred_off"])
    check(whole, "whole", edges=False)
    assert np.array_equal(start.semantic, case["sem"])            # the inputs are not modified


def test_label_propagation_edge_cases():
    """No labeled superpoint: nothing changes.  A labeled superpoint hands its label only to unlabeled neighbours whose
    predicted class agrees (and, for the extension, whose confidence exceeds 0.8); whole-scene propagation ignores
    superpoints whose predicted class has no prior or whose nearest prior is farther than 0.9 m."""
    from wsis_b200 import cluster, labels as LB
    xyz = np.array([[0, 0, 0], [0.1, 0, 0], [1, 0, 0], [1.1, 0, 0], [5, 0, 0], [5.1, 0, 0], [0.5, 0, 0]], np.float32)
    sp = np.array([0, 0, 1, 1, 2, 2, 3])
    nbrs = cluster.neighbors_from_edges(np.array([[0, 1], [1, 2], [0, 3]]), 4)
    assert nbrs == [[1, 3], [0, 2], [1], [0]]
    geo = LB.SuperpointGeometry(xyz, sp)
    assert np.allclose(geo.centre[:, 0], [0.05, 1.05, 5.05, 0.5]) and list(geo.count) == [2, 2, 2, 1]
    none = LB.SuperpointLabels([-100] * 4, [-100] * 4, np.zeros((4, 3)))
    pred, conf = np.array([3, 3, 3, 7]), np.array([0.9, 0.95, 0.5, 0.99])
    for fn in (LB.extend_label_to_neighbor, LB.propagate_label_to_neighbor):
        out = fn(none, geo, nbrs, conf, pred)
        assert np.array_equal(out.semantic, none.semantic) and not out.offset.any()
    start = LB.SuperpointLabels([3, -100, -100, -100], [11, -100, -100, -100], np.array([[0.2, 0, 0]] + [[0, 0, 0]] * 3))
    ext = LB.extend_label_to_neighbor(start, geo, nbrs, conf, pred)
    assert list(ext.semantic) == [3, 3, -100, -100] and list(ext.instance) == [11, 11, -100, -100]   # 3: other class
    assert np.allclose(ext.offset[1], [0.05 + 0.2 - 1.05, 0, 0])      # points at superpoint 0's instance centre
    assert list(LB.edge_same_instance(np.array([[0, 1], [1, 2], [0, 3]]), ext)) == [-1, 0, 0]
    prop = LB.propagate_label_to_neighbor(ext, geo, nbrs, conf, pred)  # no confidence test: 2 joins through 1
    assert list(prop.semantic) == [3, 3, 3, -100]
    whole = LB.propagate_label_to_whole_scene(start, geo, pred, np.zeros((4, 3), np.float32))
    # prior instance centre = 0.05 + 0.2 = 0.25: superpoint 1 (1.05, distance 0.8 < 0.9) joins, 2 (5.05) is too far,
    # 3 predicts a class without a prior
    assert list(whole.semantic) == [3, 3, -100, -100] and list(whole.instance) == [11, 11, -100, -100]
    assert np.allclose(whole.offset[1], 0)                            # the centroid of its own points


def test_bench_clock_sampler_summary_window_and_reasons():
    """bench.py's clock line: only samples taken inside the timed window count, throttle reasons are collected, and a
    missing nvidia-smi is reported instead of crashing."""
    sys.path.insert(0, ROOT)
    import bench
    cs = bench.ClockSampler(0)
    cs.proc = object()                                           # "running"
    mk = lambda sm, hw, pc: "0, %d, 1965, 400.0, %s, Not Active, Not Active, %s" % (sm, hw, pc)
    cs.lines = [(10.0, mk(300, "Not Active", "Not Active")),     # before the window (idle clocks)
                (20.0, mk(1965, "Not Active", "Not Active")), (20.1, mk(1950, "Not Active", "Active")),
                (20.2, mk(1965, "Not Active", "Not Active")), (30.0, mk(500, "Active", "Not Active"))]
    out = cs.summary(19.9, 20.25)
    assert out["samples"] == 3 and out["sm_mhz"] == 1965 and out["sm_max_mhz"] == 1965
    assert out["reasons"] == ["sw_power_cap"]                     # the hw_slowdown sample lies outside the window
    assert cs.summary(100.0, 101.0)["samples"] == 3               # nothing inside: falls back to the last samples
    cs.proc = None
    assert cs.summary(0, 1)["reasons"] == ["nvidia-smi unavailable"]


# ---------------------------------------------------------------------------------------------------------
# training step: loss against the reference's own MultiTaskLoss, optimizer and data-parallel plumbing
# ---------------------------------------------------------------------------------------------------------
def _loss_case(golden_dir, device="cpu"):
    g = np.load(os.path.join(golden_dir, "loss_batch2.npz"))
    pred = ("semantic_scores", "sp_semantic_scores", "pred_sp_offset_vectors", "pred_sp_occupancy", "pred_sp_ins_size",
            "sp_discriminative_feats")
    t = {k: torch.from_numpy(g[k]).to(device) for k in g.files if g[k].ndim > 0 and not k.startswith("grad_")}
    for k in pred:
        t[k].requires_grad_(True)
    return g, t, pred


@pytest.mark.parametrize("tag,epoch", [("early", 1), ("joint", 121)])
def test_vectorised_loss_matches_reference_golden(golden_dir, tag, epoch):
    """wsis_b200.train.MultiTaskLoss (one vectorised formulation for the batch) against the reference's own
    MultiTaskLoss executed by tests/golden/make_golden_loss.py (losses_3D_WSIS.py:43-230, per-scene Python loop):
    every term, the total and the gradient w.r.t. every network output."""
    from wsis_b200 import train as T
    g, t, pred = _loss_case(golden_dir)
    ret = {k: t[k] for k in pred}
    loss, parts = T.MultiTaskLoss()(T.loss_inputs(ret, t), epoch)
    loss.backward()
    assert abs(loss.item() - float(g["loss_" + tag])) < 1e-5 * abs(float(g["loss_" + tag]))
    for k, v in parts.items():
        assert abs(v.item() - float(g["%s_%s" % (k, tag)])) < 2e-5 * max(1.0, abs(float(g["%s_%s" % (k, tag)]))), k
    for k in pred:
        ref = g["grad_%s_%s" % (k, tag)] if ("grad_%s_%s" % (k, tag)) in g.files else None
        if ref is None:
            assert t[k].grad is None or float(t[k].grad.abs().max()) == 0.0
        else:
            assert np.abs(t[k].grad.numpy() - ref).max() < 1e-5 * max(np.abs(ref).max(), 1e-12), k


def test_small_tcgen05_kernels_barrier_protocol_models():
    """tools/protocol_model.py also transcribes the mbarrier protocols of ecc_messages_kernel (csrc/ecc_umma.cu) and
    wgrad_umma_kernel (csrc/wgrad_umma.cu).  The first ecc_messages protocol let thread 0 start the next operand load
    before every warp had tested the barrier for the current one: under random schedules a held-up warp finds the barrier
    two phases ahead and waits forever (on the GPU: a watchdog trap, but only with CTAs of a second stream co-resident on
    the SM).  The model must catch that protocol and pass the shipped ones."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import protocol_model as pm
    for seed in range(150):
        assert pm.EccMessagesCta(16, fixed=True).run(seed)
        assert pm.WgradCta(14, warps=4).run(seed)
        assert pm.WgradCta(3, warps=16).run(seed)
    caught = 0
    for seed in range(60):
        try:
            pm.EccMessagesCta(16, fixed=False).run(seed)
        except pm.ProtocolError:
            caught += 1
    assert caught >= 5, caught


def test_prepared_geometry_keys_match_the_network():
    """pipeline.prepare_geometry pre-builds the rulebooks under the `indice_key`s the network's convolutions ask for
    (conv.py:140-147): every sparse conv of the mirror network (3x3 / strided / inverse; the 1x1 shortcuts have no
    rulebook) must find its key in that list, and no listed key may be unused."""
    import spconv
    from wsis_b200 import pipeline
    net = pipeline.build_network(seed=1, device="cpu")
    used = {m.indice_key for m in net.modules() if isinstance(m, spconv.conv.SparseConvolution) and not m.conv1x1}
    assert used == set(pipeline.unet_rulebook_keys(pipeline.DEFAULT_MODEL_CFG["blocks"]))
