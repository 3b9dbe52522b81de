#!/usr/bin/env python
"""Offline (CPU, numpy) statistics of the Morton tiles of a synthetic scene, level by level of the U-Net: what the
conv kernel's design decisions are sized by.

  nU        distinct source rows per 128-row tile            -> row-cache capacity (kRcap = 256)
  P         entries (pairs) per tile, active offsets          -> units per tile, entries per unit
  groups    A blocks per tile if offsets with disjoint valid-slot sets share a block (greedy, densest first)
  compact   128-entry A blocks per 128 output rows if the entries of one offset are compacted over a super-tile of
            R output rows (the pair-proportional form), against the n_active blocks of the masked form

    python tools/tile_stats.py [--points 150000] [--seed 2000]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def spread3(v):
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def neighbour_map(vox):
    n = len(vox)
    key = (vox[:, 0] * 4096 + vox[:, 1]) * 4096 + vox[:, 2]
    srt = np.argsort(key)
    skey = key[srt]
    nbr = np.full((n, 27), -1, np.int64)
    k = 0
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                q = vox + np.array([dx, dy, dz])
                qk = (q[:, 0] * 4096 + q[:, 1]) * 4096 + q[:, 2]
                pos = np.minimum(np.searchsorted(skey, qk), n - 1)
                hit = (q >= 0).all(1) & (skey[pos] == qk)
                nbr[hit, k] = srt[pos[hit]]
                k += 1
    return nbr


def level_stats(vox, name, super_rows=(512, 1024)):
    n = len(vox)
    m = spread3(vox[:, 0]) << 2 | spread3(vox[:, 1]) << 1 | spread3(vox[:, 2])
    order = np.argsort(m, kind="stable")
    nbr = neighbour_map(vox)
    valid = nbr[order] >= 0
    nus, ps, nact, ngroups = [], [], [], []
    for t in range(0, n, 128):
        sub = nbr[order[t:t + 128]]
        v = sub[sub >= 0]
        nus.append(len(np.unique(v)))
        ps.append(len(v))
        vt = valid[t:t + 128]
        ks = sorted([k for k in range(27) if vt[:, k].any()], key=lambda k: -vt[:, k].sum())
        nact.append(len(ks))
        groups = []
        for k in ks:
            col = vt[:, k]
            for g in groups:
                if not (g & col).any():
                    g |= col
                    break
            else:
                groups.append(col.copy())
        ngroups.append(len(groups))
    nus, ps = np.array(nus), np.array(ps)
    line = ("%-3s rows %7d  pairs/row %5.2f | nU mean %3.0f p90 %3d p99 %3d max %3d | P/tile %4.0f | active %4.1f "
            "groups %4.1f" % (name, n, ps.sum() / n, nus.mean(), *np.percentile(nus, [90, 99]).astype(int), nus.max(),
                              ps.mean(), np.mean(nact), np.mean(ngroups)))
    for R in super_rows:
        blocks = 0
        for t in range(0, n, R):
            blocks += int(np.ceil(valid[t:t + R].sum(0) / 128.0).sum())
        line += " | compact R=%d: %.1f blocks/128 rows" % (R, blocks / (n / 128.0))
    print(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--seed", type=int, default=2000)
    args = ap.parse_args()
    from wsis_b200 import synthetic
    sc = synthetic.make_scene(args.seed, n_points=args.points)
    vox = np.unique(sc["locs"], axis=0)
    for lvl in range(1, 6):
        level_stats(vox, "L%d" % lvl)
        vox = np.unique(vox // 2, axis=0)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(40), indexing="ij"), -1).reshape(-1, 3)
    level_stats(g, "blob")


if __name__ == "__main__":
    main()
