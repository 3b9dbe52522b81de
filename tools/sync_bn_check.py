#!/usr/bin/env python
"""Run under torchrun with >= 2 GPUs: the in-kernel statistics all-reduce over NVLink peer memory (csrc/train.cu
bn_sync_kernel) against NCCL dist.all_reduce and against torch.nn.SyncBatchNorm: forward output, running statistics, input
/ gamma / beta gradients; every rank must hold bitwise identical statistics.  Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from wsis_b200 import train as T  # noqa: E402

res = {"world": world, "cases": []}
for C, N in ((32, 5000 + 777 * rank), (64, 20000 + 13 * rank), (224, 900 + rank), (20, 3001)):
    torch.manual_seed(100 + rank + C)
    x = torch.randn(N, C, device="cuda") * (1 + rank) + 0.3 * rank
    g = torch.randn(N, C, device="cuda")
    outs = {}
    for transport in ("peer", "nccl", "torch"):
        bn = torch.nn.BatchNorm1d(C, eps=1e-4, momentum=0.1).cuda().train()
        with torch.no_grad():
            torch.manual_seed(C)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.5, 0.5)
        xi = x.clone().requires_grad_(True)
        if transport == "torch":
            sbn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(bn)
            y = torch.relu(sbn(xi))
            bn = sbn
        else:
            T.SYNC_BN_TRANSPORT = transport
            for _ in range(3 if transport == "peer" else 1):      # several calls: sequence numbers / parity slots cycle
                bn.running_mean.zero_(), bn.running_var.fill_(1.0)
                xi.grad = None
                y = T.batch_norm_train(xi, bn, True)
        y.backward(g)
        outs[transport] = [y.detach(), bn.running_mean.clone(), bn.running_var.clone(), xi.grad.clone(), bn.weight.grad.clone(),
                           bn.bias.grad.clone()]
    assert T._PEER is not None or T._PEER_FAILED, "peer transport was never initialised"
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))  # noqa: E731
    case = {"C": C, "N_rank": N, "peer_vs_nccl": max(rel(a, b) for a, b in zip(outs["peer"], outs["nccl"])),
            "peer_vs_torch_syncbn": max(rel(a, b) for a, b in zip(outs["peer"], outs["torch"]))}
    # every rank must hold the same running statistics bit for bit
    rm = outs["peer"][1].clone()
    ref = rm.clone()
    dist.broadcast(ref, 0)
    case["ranks_bitwise_equal"] = bool(torch.equal(rm, ref))
    flags = torch.tensor([case["peer_vs_nccl"], case["peer_vs_torch_syncbn"], 0.0 if case["ranks_bitwise_equal"] else 1.0],
                         device="cuda", dtype=torch.float64)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    case["peer_vs_nccl"], case["peer_vs_torch_syncbn"], case["ranks_bitwise_equal"] = float(flags[0]), float(flags[1]), flags[2].item() == 0.0
    res["cases"].append(case)
res["peer_transport_active"] = not T._PEER_FAILED
ok = all(c["peer_vs_nccl"] < 1e-6 and c["peer_vs_torch_syncbn"] < 3e-5 and c["ranks_bitwise_equal"] for c in res["cases"])
res["ok"] = bool(ok and res["peer_transport_active"])
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
