import sys, os
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from wsis_b200 import ops as W, synthetic
import pointgroup_ops
dev="cuda"
batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=150000) for i in range(4)])
locs, _, _ = pointgroup_ops.voxelization_idx(batch["locs"].to(dev), 4, 4)
coords, shape = locs.int(), batch["spatial_shape"]
for lvl in range(1,6):
    rb = W.rulebook_subm(coords, shape, 3, 1, batch_size=4)
    t = rb.tiles_out()
    meta = t.meta.cpu().numpy()
    nU, P = meta[:,1], meta[:,3]
    nact = np.array([bin(int(m)&0xffffffff).count("1") for m in meta[:,2]])
    print("level", lvl, "rows", coords.shape[0], "tiles", t.num_tiles, "nU pct 50/90/99/max", np.percentile(nU,[50,90,99]).tolist(), nU.max(), "frac>256 %.3f"%((nU>256).mean()), "frac>384 %.3f"%((nU>384).mean()), "P/tile %.0f"%P.mean(), "nact %.1f"%nact.mean())
    rbc, shape = W.rulebook_conv(coords, shape, 2, 2, 0, 1, batch_size=4)
    # also strided/inverse tile stats
    for name, tt in (("down", rbc.tiles_out()), ("up", rbc.tiles_in())):
        m = tt.meta.cpu().numpy()
        na_ = np.array([bin(int(x)&0xffffffff).count("1") for x in m[:,2]])
        print("   ", name, "tiles", tt.num_tiles, "nU 50/max", np.percentile(m[:,1],50), m[:,1].max(), "nact %.1f"%na_.mean(), "P/tile %.0f"%m[:,3].mean())
    coords = rbc.out_coords
