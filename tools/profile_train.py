#!/usr/bin/env python
"""Kernel-time table of one training step (torch profiler, CUDA activity): which kernels the backward spends its time in.
    python tools/profile_train.py [--points 150000] [--scenes 4] [--mode train|infer] > gpurun_out/train_kernels.txt"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=150000)
ap.add_argument("--scenes", type=int, default=4)
ap.add_argument("--mode", default="train")
ap.add_argument("--rows", type=int, default=45)
args = ap.parse_args()
from wsis_b200 import pipeline, synthetic, train as T  # noqa: E402

batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=args.points) for i in range(args.scenes)], with_labels=True)
db = pipeline.to_device(batch)[0]
if args.mode == "train":
    net = pipeline.build_network(seed=123, device="cuda").train()
    step = T.TrainStep(net)
    fn = lambda: step(db)  # noqa: E731
else:
    net = pipeline.build_network(seed=123, device="cuda").eval()

    def fn():
        with torch.no_grad():
            pipeline.forward_batch(net, db)
for _ in range(3):
    fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=90))
