#!/usr/bin/env python
"""Random-walk label propagation (scannetv2_dataset.py:664-735): the device kernels (csrc/affinity.cu rw_*) timed with
CUDA events against the reference's dense float64 numpy formulation (oracle/oracle.py, bit-equal to the reference's own
method) on the superpoint graph of a full 150k-point scene, iterations 0..3."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from wsis_b200 import ops as W  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "rw_scene0.npz"))
S, e = int(g["S"]), g["edges"].astype(np.int64)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
eu, ev, aff = cu(e[:, 0]), cu(e[:, 1]), cu(g["aff"])
seed, pred, conf = cu(g["seed_label"].astype(np.int64)), cu(g["pred"].astype(np.int64)), cu(g["conf"])
useg, vseg = W.SegmentIndex(eu, S), W.SegmentIndex(ev, S)
adj = np.zeros((S, S))
adj[e[:, 0], e[:, 1]] = 1
A = orc.dense_affinity(e[:, 0], e[:, 1], g["aff"], S)
for it in (0, 1, 3):
    ts = []
    for rep in range(7):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        pseudo, score = W.random_walk(eu, ev, aff, seed, pred, conf, 20, it, useg=useg, vseg=vseg)
        t.record()
        torch.cuda.synchronize()
        if rep >= 2:
            ts.append(s.elapsed_time(t))
    t0 = time.perf_counter()
    final, sc = orc.weak_label_propagation(g["seed_label"].astype(np.int64), adj, g["conf"], g["pred"].astype(np.int64), A, it)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(pseudo.cpu().numpy(), final.astype(np.int32)))
    print(json.dumps({"S": S, "edges": int(len(e)), "seeds": int((g["seed_label"] != -100).sum()), "iterations": it,
                      "gpu_ms": round(float(np.median(ts)), 3), "cpu_numpy_s": round(cpu_s, 3),
                      "speedup": round(cpu_s * 1e3 / float(np.median(ts)), 1), "labels_identical": same,
                      "labelled": int((final != -100).sum()), "cores": os.cpu_count()}), flush=True)
