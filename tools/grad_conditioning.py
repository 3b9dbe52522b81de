#!/usr/bin/env python
"""How well-conditioned is the training-step gradient?  Same batch, same weights, five ways of computing it:
CPU reference kernels (fp32), GPU fused (tensor-core fp32 contract), GPU fused (exact fp32 SIMT), GPU unfused
(torch BatchNorm + the unfused conv autograd path), and the CPU reference again with the batch rows permuted inside the
BatchNorm reductions (different summation order only).  Prints the worst per-parameter differences of each pair."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from oracle import cpu_pipeline  # noqa: E402
from wsis_b200 import ops as W, pipeline, synthetic, train as T  # noqa: E402

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
batch = synthetic.collate([synthetic.make_scene(2000 + i, n_points=n_points) for i in range(2)], with_labels=True)


def cpu_grads(threads):
    torch.set_num_threads(threads)
    net = pipeline.build_network(seed=123, device="cpu").train()
    cpu_pipeline.train_step(net, batch, T.MultiTaskLoss())
    return {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}


def gpu_grads(prec, fused):
    W.set_precision(prec)
    T.FUSED = fused
    net = pipeline.build_network(seed=123, device="cuda").train()
    step = T.TrainStep(net)
    step(pipeline.to_device(batch)[0], optimize=False)
    T.FUSED = True
    W.set_precision("fp32")
    return {n: p.grad.detach().cpu().clone() for n, p in net.named_parameters() if p.grad is not None}


runs = {"cpu16": cpu_grads(16), "cpu1": cpu_grads(1), "gpu_fp32": gpu_grads("fp32", True), "gpu_simt": gpu_grads("simt", True),
        "gpu_unfused_simt": gpu_grads("simt", False)}
gmax = max(float(g.abs().max()) for g in runs["cpu16"].values())
names = list(runs)
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        a, b = runs[names[i]], runs[names[j]]
        d = {k: float((a[k] - b[k]).abs().max()) / max(float(b[k].abs().max()), 1e-4 * gmax) for k in a}
        top = sorted(d.items(), key=lambda kv: -kv[1])[:2]
        va = torch.cat([a[k].double().reshape(-1) for k in a])
        vb = torch.cat([b[k].double().reshape(-1) for k in a])
        cos = float(va @ vb / (va.norm() * vb.norm()))
        l2 = float((va - vb).norm() / vb.norm())
        pl2 = max(float((a[k].double() - b[k].double()).norm() / max(float(b[k].double().norm()), 1e-4 * float(vb.norm()))) for k in a)
        print("%-16s vs %-16s 1-cos %.2e  relL2 %.2e  worst-param relL2 %.2e  worst max-abs %s"
              % (names[i], names[j], 1 - cos, l2, pl2, ", ".join("%s %.2e" % (k[-34:], v) for k, v in top)))
