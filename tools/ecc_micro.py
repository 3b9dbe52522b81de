#!/usr/bin/env python
"""ECC-GRU microbenchmark on a 4-scene batch's superpoint graph: filter-free tensor-core path (csrc/ecc_umma.cu) vs the
streamed-filter path (library GEMMs materialise [E,1024], csrc/ecc.cu streams it 7 times)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from wsis_b200 import model as M, ops as W, synthetic  # noqa: E402

batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=150000) for i in range(4)])
S, E = batch["num_superpoints"], batch["ecc_edge_index"].shape[1]
torch.manual_seed(0)
fnet = M.create_fnet([13, 32, 128, 64, 32 * 32], True, True, 2)
cell = M.GRUCellEx(32, 32, bias=True, layernorm=True, ingate=True)
mod = M.RNNGraphConvModule(cell, fnet, 32, nrepeats=7, cat_all=True).cuda().eval()
mod.set_info(M.GraphInfo(batch["ecc_edge_index"].cuda(), batch["ecc_edgefeats"].cuda()))
hx = torch.randn(S, 32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for fused in (True, False):
    W.ECC_FUSED_FILTERS = fused
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        with torch.no_grad():
            out = mod(hx)
        e.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(s.elapsed_time(e))
    print(json.dumps({"superpoints": S, "edges": E, "filter_free": fused, "ms_7_steps_incl_filter_net": round(float(np.median(ts)), 3),
                      "materialised_filter_MB": 0 if fused else round(E * 4096 / 1e6, 1)}), flush=True)
