#!/usr/bin/env python
"""Instance clustering (test_scannetv2.py:281-455) on a full 150k-point scene: the device path (csrc/cluster.cu) timed with
CUDA events against the host restatement (wsis_b200/cluster.py, itself ~28x faster than the reference's N-point-mask
formulation), on the network's own outputs."""
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
from wsis_b200 import cluster, pipeline, synthetic  # noqa: E402

sc = synthetic.make_scene(2000, n_points=150000)
batch = synthetic.collate([sc])
net = pipeline.build_network(seed=123, device="cuda").eval()
db = pipeline.to_device(batch)[0]
with torch.no_grad():
    ret, aux = pipeline.forward_batch(net, db)
S = sc["num_superpoints"]
sem = ret["sp_semantic_scores"].max(1)[1]
xyz = torch.from_numpy(sc["xyz"]).cuda()
edges = torch.from_numpy(sc["edges"]).cuda()
ts = []
for rep in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    csr = cluster.neighbors_csr_device(edges, S)
    conf, label, point_inst, _ = cluster.clustering_in_graph_device(xyz, db["superpoint"], csr, sem, ret["pred_sp_offset_vectors"],
                                                                  ret["pred_sp_occupancy"], ret["pred_sp_ins_size"], num_superpoints=S)
    e.record()
    torch.cuda.synchronize()
    if rep >= 1:
        ts.append(s.elapsed_time(e))
t0 = time.perf_counter()
nbrs = cluster.neighbors_from_edges(sc["edges"], S)
hconf, hlabel, hmasks = cluster.clustering_in_graph(sc["xyz"], sc["superpoint"], nbrs, sem.cpu().numpy(),
                                                    ret["pred_sp_offset_vectors"].cpu().numpy(), ret["pred_sp_occupancy"].cpu().numpy(),
                                                    ret["pred_sp_ins_size"].cpu().numpy(), dense=False)
host_s = time.perf_counter() - t0
host_inst = hmasks
print(json.dumps({"points": 150000, "superpoints": S, "instances": int(conf.shape[0]), "device_ms": round(float(np.median(ts)), 3),
                  "host_numpy_s": round(host_s, 3), "masks_identical": bool(np.array_equal(point_inst.cpu().numpy(), host_inst))}))
