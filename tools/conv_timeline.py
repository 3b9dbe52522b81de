#!/usr/bin/env python
"""Timeline of CTA 0 of the tensor-core conv kernel (wsis_conv_debug_timeline): which role waits for which, when.
    python tools/conv_timeline.py [--shape 32x32] [--precision fp32] > gpurun_out/timeline.txt"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="32x32")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--scenes", type=int, default=4)
    args = ap.parse_args()
    from wsis_b200 import ops as W, synthetic
    from wsis_b200._lib import lib
    import pointgroup_ops
    dev = "cuda"
    batch = synthetic.collate([synthetic.make_scene(2000 + i) for i in range(args.scenes)])
    locs, _, _ = pointgroup_ops.voxelization_idx(batch["locs"].to(dev), args.scenes, 4)
    coords, shape = locs.int(), batch["spatial_shape"]
    rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=args.scenes)
    tiles = rb.tiles_out()
    N = coords.shape[0]
    cin, cout = (int(x) for x in args.shape.split("x"))
    x = torch.rand((N, cin), device=dev) - 0.5
    w = (torch.rand((27, cin, cout), device=dev) - 0.5) / cin ** 0.5
    packed = W.PackedWeights()
    for _ in range(2):
        W.sparse_conv(x, w, rb.nbr_in, N, 1, packed=packed, precision=args.precision, tiles=tiles)
    cap = 32 * 512
    buf = torch.zeros((2 * cap,), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib().call("wsis_conv_debug_timeline", ctypes.c_void_p(buf.data_ptr()), cap)
    W.sparse_conv(x, w, rb.nbr_in, N, 1, packed=packed, precision=args.precision, tiles=tiles)
    torch.cuda.synchronize()
    lib().call("wsis_conv_debug_timeline", None, 0)
    b = buf.cpu().numpy()
    ev = b.reshape(cap, 2)
    ev = ev[ev[:, 0] > 0]
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    t0 = ev[0, 0]
    names = {0: "epi", 1: "gather", 24: "wprod"}
    evn = {0: {0: "acc_full", 1: "done"}, 1: {0: "rec_ready", 1: "loads_issued", 2: "rc_free", 3: "rc_full"},
           2: {0: "prefetched", 1: "stage_free", 2: "a_full"}, 16: {0: "a_ready", 1: "w_ready", 2: "issued"},
           24: {0: "w_free"}}
    for t, code in ev:
        role, it, e, unit = (code >> 24) & 0xff, (code >> 16) & 0xff, (code >> 12) & 0xf, code & 0xfff
        base = 2 if 2 <= role < 16 else 16 if 16 <= role < 24 else role
        rn = names.get(role, ("build%d" % (role - 2)) if base == 2 else ("issue%d" % (role - 16)))
        print("%9d cyc %-8s tile %d  %-12s unit %d" % (t - t0, rn, it, evn[base].get(e, str(e)), unit))


if __name__ == "__main__":
    main()
