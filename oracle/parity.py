"""TEST / BASELINE INFRASTRUCTURE ONLY -- full-size parity of the device path against the reference's own CPU
kernels (oracle/cpu_pipeline.forward with kind == "reference"): the nine rulebooks of the U-Net
(modules/lib/spconv/spconv/conv.py:149-152 -> indice_dict keys subm1-5, spconv1-4) as sorted pair sets in coordinate
space, the voxelization maps, the U-Net output and every entry of the result dict of backbone_3D_WSIS.py:164-255.

Used by tests/test_gpu_parity.py and by bench.py's cpu_baseline leg (the `parity` object of the bench line)."""
import numpy as np
import torch

KEYS = ["subm%d" % i for i in range(1, 6)] + ["spconv%d" % i for i in range(1, 5)]


def _ckey(coords):
    c = np.asarray(coords).astype(np.int64)
    return (c[:, 0] << 48) | (c[:, 1] << 32) | (c[:, 2] << 16) | c[:, 3]


def canonical_pairs(in_coords, out_coords, pairs, num):
    """Reference-format rulebook (pairs int32[K,2,N], num int32[K]) -> int64[P,3] rows (k, in coordinate key, out
    coordinate key), sorted: equal for two builders iff their rulebooks are equal as pair sets (the reference's GPU
    order is atomics-dependent, indice.cu.h:57,202, so order is not part of the contract)."""
    pairs, num = np.asarray(pairs), np.asarray(num)
    ik, ok = _ckey(in_coords), _ckey(out_coords)
    rows = []
    for k in range(pairs.shape[0]):
        n = int(num[k])
        if n:
            rows.append(np.stack([np.full(n, k, np.int64), ik[pairs[k, 0, :n]], ok[pairs[k, 1, :n]]], 1))
    if not rows:
        return np.zeros((0, 3), np.int64)
    r = np.concatenate(rows)
    return r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if a.size else 0.0


def device_rulebooks(aux):
    """{key: (in coords, out coords, pairs, num)} as numpy arrays from the device forward's indice_dict."""
    out = {}
    d = aux["input"].indice_dict
    for key in KEYS:
        if key not in d:
            continue
        outids, indices, pairs, num, _ = d[key]
        if hasattr(pairs, "tensors"):  # inference keeps the reference-format tensors lazy
            pairs, num = pairs.tensors()
        out[key] = (indices.cpu().numpy(), outids.cpu().numpy(), pairs.cpu().numpy(), num.cpu().numpy())
    return out


def compare(ret, aux, cpu_ret, cpu_keep):
    """-> dict: rulebooks_equal, per-key pair counts, voxelization_equal, max relative error (max-abs normalised) of the
    U-Net output and of every result entry."""
    res = {"rulebooks": {}, "outputs": {}}
    dev = device_rulebooks(aux)
    level_coords = {}
    ok_all = True
    for key in KEYS:
        outids, pairs, num, _ = cpu_keep["rulebooks"][key]
        lvl = int(key[-1])
        if key.startswith("subm"):
            level_coords[lvl] = outids.numpy()
            cin = cout = outids.numpy()
        else:
            cin, cout = level_coords[lvl], outids.numpy()
            level_coords[lvl + 1] = cout
        ref = canonical_pairs(cin, cout, pairs.numpy(), num.numpy())
        if key not in dev:
            res["rulebooks"][key] = {"pairs": int(ref.shape[0]), "equal": False, "missing": True}
            ok_all = False
            continue
        got = canonical_pairs(*dev[key])
        same_out = np.array_equal(np.sort(_ckey(dev[key][1])), np.sort(_ckey(cout)))
        eq = bool(got.shape == ref.shape and np.array_equal(got, ref) and same_out)
        ok_all &= eq
        res["rulebooks"][key] = {"pairs": int(ref.shape[0]), "equal": eq}
    res["rulebooks_equal"] = bool(ok_all)
    res["voxelization_equal"] = bool(
        np.array_equal(aux["voxel_locs"].cpu().numpy(), cpu_keep["voxel_locs"].numpy())
        and np.array_equal(aux["p2v_map"].cpu().numpy(), cpu_keep["p2v"].numpy())
        and np.array_equal(aux["v2p_map"].cpu().numpy(), cpu_keep["v2p"].numpy()))
    if aux.get("unet_features") is not None:
        res["outputs"]["unet_features"] = rel(aux["unet_features"].float().cpu().numpy(), cpu_keep["unet_out"].numpy())
    for k, v in cpu_ret.items():
        res["outputs"][k] = rel(ret[k].float().cpu().numpy(), v.numpy())
    res["max_rel"] = max(res["outputs"].values())
    return res
