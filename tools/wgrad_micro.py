#!/usr/bin/env python
"""Weight-gradient microbenchmark: dW of a submanifold 3x3x3 conv on the bench scenes' voxel sets, tensor-core kernel
(csrc/wgrad_umma.cu) vs the SIMT kernel, CUDA-event timing with an L2 flush between iterations.
    python tools/wgrad_micro.py [--level 1] [--shapes 32x32,64x64] [--debug 0,1,2,4]"""
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

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=4)
ap.add_argument("--points", type=int, default=150000)
ap.add_argument("--level", type=int, default=1)
ap.add_argument("--shapes", default="32x32,64x64")
ap.add_argument("--debug", default="0")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--simt", action="store_true")
args = ap.parse_args()
from wsis_b200 import ops as W, synthetic  # noqa: E402
import pointgroup_ops  # noqa: E402

dev = "cuda"
batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=args.points) for i in range(args.scenes)])
locs, _, _ = pointgroup_ops.voxelization_idx(batch["locs"].to(dev), args.scenes, 4)
coords, shape, bs = locs.int(), batch["spatial_shape"], args.scenes
for _ in range(args.level - 1):
    rbc, shape = W.rulebook_conv(coords, shape, 2, 2, 0, 1, batch_size=bs)
    coords = rbc.out_coords
N = coords.shape[0]
rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=bs)
P = int((rb.nbr_in >= 0).sum().item())
order = rb.order_hint("out")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for shp in args.shapes.split(","):
    cin, cout = (int(x) for x in shp.split("x"))
    x = torch.rand((N, cin), device=dev, generator=g) - 0.5
    go = torch.rand((N, cout), device=dev, generator=g) - 0.5
    scale, shift = torch.rand(cin, device=dev, generator=g) + 0.5, torch.rand(cin, device=dev, generator=g) - 0.5
    for dbg in args.debug.split(","):
        os.environ["WSIS_WGRAD_DEBUG"] = dbg
        for prec in (["fp32", "bf16"] + (["simt"] if args.simt and dbg == "0" else [])):
            ts = []
            for it in range(args.iters + 2):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                W.sparse_conv_wgrad(x, rb.nbr_in, N, 1, go, 27, cin, cout, prologue=(scale, shift, 1), order=order, precision=prec)
                e.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(s.elapsed_time(e) * 1e-3)
            t = float(np.median(ts))
            b = 4 * N * (cin + cout) + 8 * P + 4 * 27 * cin * cout
            print(json.dumps({"voxels": N, "pairs": P, "cin": cin, "cout": cout, "precision": prec, "debug": int(dbg),
                              "us": round(t * 1e6, 1), "alg_MB": round(b / 1e6, 1), "GBps": round(b / t / 1e9, 1),
                              "useful_tflops": round(2.0 * P * cin * cout / t / 1e12, 2)}), flush=True)
