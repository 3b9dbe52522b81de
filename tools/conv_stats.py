#!/usr/bin/env python
"""Where the roles of the tensor-core conv kernel wait (wsis_conv_debug_stats): CTA 0 of one launch, cycles per role.

    python tools/conv_stats.py [--level 2] [--shape 64x64] [--precision fp32]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROLES = [("epilogue", ["acc_full"])] * 4 + [("gather", ["record", "rc_free"])] * 4 + \
        [("build", ["record", "rc_full", "slot_free", "tmem_store"])] * 8 + [("issue", ["record", "acc_free", "stage_full", "mma_issue"])] * 4 + \
        [("records", ["buf_free"])] + [("weights", ["stage_free"])] * 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--shape", default="64x64")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--timeline", type=int, nargs=2, default=None, help="print CTA 0's events between two cycle counts")
    args = ap.parse_args()
    from wsis_b200 import ops as W
    from wsis_b200 import synthetic
    from wsis_b200._lib import lib
    import pointgroup_ops
    dev = "cuda"
    batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=args.points) for i in range(args.scenes)])
    locs, _, _ = pointgroup_ops.voxelization_idx(batch["locs"].to(dev), args.scenes, 4)
    coords, shape, bs = locs.int(), batch["spatial_shape"], args.scenes
    for _ in range(args.level - 1):
        rbc, shape = W.rulebook_conv(coords, shape, 2, 2, 0, 1, batch_size=bs)
        coords = rbc.out_coords
    N = coords.shape[0]
    rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=bs)
    tiles = rb.tiles_out()
    cin, cout = (int(x) for x in args.shape.split("x"))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand((N, cin), device=dev, generator=g) - 0.5
    w = (torch.rand((27, cin, cout), device=dev, generator=g) - 0.5) / cin ** 0.5
    packed = W.PackedWeights()
    W.sparse_conv(x, w, rb.nbr_in, N, 1, packed=packed, precision=args.precision, tiles=tiles)
    buf = torch.zeros(256 + 24 * 512, dtype=torch.int64, device=dev)
    lib().call("wsis_conv_debug_stats", ctypes.c_void_p(buf.data_ptr()))
    W.sparse_conv(x, w, rb.nbr_in, N, 1, packed=packed, precision=args.precision, tiles=tiles)
    torch.cuda.synchronize()
    lib().call("wsis_conv_debug_stats", None)
    raw = buf.cpu().numpy()
    st = raw[:23 * 8].reshape(23, 8)
    tiles_cta0 = -(-tiles.num_tiles // 148)
    meta = tiles.meta.cpu().numpy()
    units = int(sum(int(m) for m in meta[0::148, 2])) * (-(-cin // (32 if args.precision == "fp32" else 64)))
    out = {"shape": args.shape, "level": args.level, "precision": args.precision, "tiles_cta0": tiles_cta0,
           "units_cta0": units, "roles": []}
    for wi, (name, waits) in enumerate(ROLES):
        tot = int(st[wi, 0])
        if tot == 0:
            continue
        out["roles"].append({"warp": wi, "role": name, "cycles": tot, "cyc_per_unit": round(tot / max(units, 1), 1),
                             "waits": {k: round(int(st[wi, 1 + i]) / tot, 3) for i, k in enumerate(waits)}})
    print(json.dumps(out, indent=None))
    if args.timeline:
        names = {1: "epi: acc full", 2: "epi: acc released", 3: "gather: record ready", 4: "gather: pass done",
                 5: "build: row cache ready", 6: "build: slots free (probe passed)", 7: "build: stage full",
                 8: "issue: acc free", 9: "issue: stage full seen", 10: "issue: stage committed", 11: "issue: tile committed",
                 12: "records: buffer free"}
        ev = []
        for w in range(24):
            for x in raw[256 + w * 512:256 + (w + 1) * 512]:
                if x:
                    ev.append((int(x) >> 8, w, int(x) & 0xff))
        ev.sort()
        lo, hi = args.timeline
        for t, w, c in ev:
            if lo <= t <= hi:
                print("%9d  warp %2d %-9s %s" % (t, w, ROLES[w][0] if w < len(ROLES) else "?", names.get(c, c)), file=sys.stderr)
    for r in out["roles"]:
        print("%2d %-9s %9d cyc  %7.1f/unit  %s" % (r["warp"], r["role"], r["cycles"], r["cyc_per_unit"], r["waits"]), file=sys.stderr)


if __name__ == "__main__":
    main()
