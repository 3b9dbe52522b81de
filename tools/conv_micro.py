#!/usr/bin/env python
"""Sparse-conv microbenchmark (BASELINE.json configs[4]): submanifold 3x3x3 on synthetic scenes, one layer shape at
a time, CUDA-event timing with an L2 flush between iterations, algorithmic bytes per SURVEY.md 8(d).

    python tools/conv_micro.py [--scenes 4] [--points 150000] [--shapes 32x32,64x64] [--precision fp32] [--iters 5]
                               [--order morton|identity] [--shell]

Also the target of the `ncu --set full` captures under profiles/ (one launch of one layer)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def occupancy_set(n_voxels, density, seed=0):
    """~n_voxels active voxels: density in (0,1] = i.i.d. occupancy of a cube; density == "surface" = an axis-aligned
    box shell (every voxel has in-plane neighbours only, the scan-like case)."""
    rng = np.random.default_rng(seed)
    if density == "surface":
        L = max(8, int(round((n_voxels / 6.0) ** 0.5)))
        a, b = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
        a, b = a.ravel(), b.ravel()
        z0, z1 = np.zeros_like(a), np.full_like(a, L - 1)
        faces = [np.stack(f, 1) for f in ((a, b, z0), (a, b, z1), (a, z0, b), (a, z1, b), (z0, a, b), (z1, a, b))]
        c = np.unique(np.concatenate(faces), axis=0)
        shape = [L, L, L]
    else:
        L = max(8, int(np.ceil((n_voxels / density) ** (1.0 / 3.0))))
        flat = rng.choice(L ** 3, size=min(n_voxels, L ** 3), replace=False)
        c = np.stack([flat // (L * L), (flat // L) % L, flat % L], 1)
        shape = [L, L, L]
    c = c[rng.permutation(len(c))]
    return np.concatenate([np.zeros((len(c), 1), np.int64), c], 1).astype(np.int32), shape


def sweep(args):
    """BASELINE.json configs[4]: submanifold 3x3x3, C = 16..64, 50k..2M active voxels, varying occupancy density."""
    from wsis_b200 import ops as W
    dev, peak = "cuda", 6551.4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for n_vox in [int(x) for x in args.sweep_voxels.split(",")]:
        for dens in args.sweep_density.split(","):
            density = dens if dens == "surface" else float(dens)
            c, shape = occupancy_set(n_vox, density)
            if max(shape) > 65535:
                continue
            coords = torch.from_numpy(c).to(dev)
            N = coords.shape[0]
            rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=1)
            P = int((rb.nbr_in >= 0).sum().item())
            tiles = rb.tiles_out()
            for C in [int(x) for x in args.sweep_channels.split(",")]:
                x = torch.rand((N, C), device=dev, generator=g) - 0.5
                w = (torch.rand((27, C, C), device=dev, generator=g) - 0.5) / C ** 0.5
                packed = W.PackedWeights()
                ts = []
                for it in range(args.iters + 2):
                    flush.zero_()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    W.sparse_conv(x, w, rb.nbr_in, N, 1, packed=packed, precision=args.precision, tiles=tiles)
                    e.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(s.elapsed_time(e) * 1e-3)
                t = float(np.median(ts))
                b = 4 * N * 2 * C + 8 * P + 4 * 27 * C * C
                print(json.dumps({"voxels": N, "density": dens, "grid": shape[0], "pairs_per_voxel": round(P / N, 2),
                                  "C": C, "precision": args.precision, "us": round(t * 1e6, 1), "alg_MB": round(b / 1e6, 1),
                                  "GBps": round(b / t / 1e9, 1), "hbm_frac": round(b / t / 1e9 / peak, 4),
                                  "useful_tflops": round(2.0 * P * C * C / t / 1e12, 2)}), flush=True)
            del rb, tiles, coords


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true", help="BASELINE.json configs[4] sweep (one JSON line per case)")
    ap.add_argument("--sweep-voxels", default="50000,150000,500000,1000000,2000000")
    ap.add_argument("--sweep-density", default="0.01,0.1,0.3,surface,1.0",
                    help="occupancy of a cube (1.0 = solid blob, ~27 pairs/voxel; 0.01 = random scatter, ~1.3) or `surface`")
    ap.add_argument("--sweep-channels", default="16,32,48,64")
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--shapes", default="32x32,64x64,64x32,128x128")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--order", default="morton", choices=["morton", "identity"])
    ap.add_argument("--shell", action="store_true", help="the 150k-voxel floor+wall shell of SURVEY.md 6 instead of scenes")
    ap.add_argument("--level", type=int, default=1, help="UNet level whose voxel set is used (1 = 2 cm, 2 = 4 cm, ...)")
    args = ap.parse_args()
    if args.sweep:
        return sweep(args)

    from wsis_b200 import ops as W
    from wsis_b200 import synthetic
    import pointgroup_ops
    dev = "cuda"
    if args.shell:
        c, shape = synthetic.make_shell()
        coords = torch.from_numpy(c.astype(np.int32)).to(dev)
        bs = 1
    else:
        batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=args.points) for i in range(args.scenes)])
        locs, _, _ = pointgroup_ops.voxelization_idx(batch["locs"].to(dev), args.scenes, 4)
        coords, shape, bs = locs.int(), batch["spatial_shape"], args.scenes
    for _ in range(args.level - 1):
        rbc, shape = W.rulebook_conv(coords, shape, 2, 2, 0, 1, batch_size=bs)
        coords = rbc.out_coords
    N = coords.shape[0]
    rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=bs)
    P = int((rb.nbr_in >= 0).sum().item())
    tiles = rb.tiles_out() if args.order == "morton" else W.TileMap(rb.nbr_in, N, 1, W.identity_order(N, coords.device))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 6551.4
    out = {"voxels": N, "pairs": P, "pairs_per_voxel": round(P / max(N, 1), 2), "order": args.order,
           "precision": args.precision, "layers": []}
    g = torch.Generator(device=dev).manual_seed(0)
    for shp in args.shapes.split(","):
        cin, cout = (int(x) for x in shp.split("x"))
        x = torch.rand((N, cin), device=dev, generator=g) - 0.5
        w = (torch.rand((27, cin, cout), device=dev, generator=g) - 0.5) / cin ** 0.5
        scale, shift = torch.rand(cin, device=dev, generator=g) + 0.5, torch.rand(cin, device=dev, generator=g) - 0.5
        res = torch.rand((N, cout), device=dev, generator=g)
        packed = W.PackedWeights()
        ts = []
        for it in range(args.iters + 2):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            W.sparse_conv(x, w, rb.nbr_in, N, 1, prologue=(scale, shift, 1), residual=res, packed=packed,
                          precision=args.precision, tiles=tiles)
            e.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(s.elapsed_time(e) * 1e-3)
        t = float(np.median(ts))
        # e(N Cin + N Cout) + 2 idx P + e K Cin Cout  (+ e N Cout for the fused residual read)
        b = 4 * N * (cin + cout) + 8 * P + 4 * 27 * cin * cout + 4 * N * cout
        out["layers"].append({"cin": cin, "cout": cout, "us": round(t * 1e6, 1), "alg_MB": round(b / 1e6, 1),
                              "GBps": round(b / t / 1e9, 1), "hbm_frac": round(b / t / 1e9 / peak, 4),
                              "useful_tflops": round(2.0 * P * cin * cout / t / 1e12, 1),
                              "us_per_tile_per_cta": round(t * 1e6 / max(1, -(-tiles.num_tiles // 148)), 2)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
