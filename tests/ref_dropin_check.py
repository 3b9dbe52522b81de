"""Run by tests/test_cpu.py::test_reference_model_files_run_on_the_dropin in a subprocess (it rewires sys.modules).

Imports the reference's OWN model files -- /root/reference/modules/model/{backbone_3D_WSIS,sparse_unet3d,graphnet,
spg_modules}.py, unmodified -- on top of THIS repository's `spconv` / `pointgroup_ops` packages (the drop-in boundary,
SURVEY.md 8b), builds `Network(cfg)` under torch.manual_seed(123) and compares its state_dict, key by key and bit by
bit, with the host mirror wsis_b200.model.Network built under the same seed.  Libraries the reference imports but that
are neither in its tree nor installed here (torch_scatter, torch_geometric, func_helper's htree/treelib) are stubbed
exactly as tests/golden/make_golden.py documents; construction uses none of them.  Prints one JSON line."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WSIS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "3d-wsis_b200"))       # spconv, pointgroup_ops, wsis_b200 = the drop-in
sys.path.insert(0, os.path.join(REF, "modules/model"))

ts = types.ModuleType("torch_scatter")
ts.scatter = ts.scatter_add = ts.scatter_mean = ts.scatter_max = ts.scatter_min = None
sys.modules["torch_scatter"] = ts
tg, tgnn, tgconv, tginits = (types.ModuleType(n) for n in ("torch_geometric", "torch_geometric.nn",
                                                           "torch_geometric.nn.conv", "torch_geometric.nn.inits"))


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", **kw):
        super().__init__()
        self.aggr, self.flow = aggr, flow


tgconv.MessagePassing = MessagePassing
tginits.uniform = lambda size, tensor: None if tensor is None else tensor.data.uniform_(-1 / size ** 0.5, 1 / size ** 0.5)
tg.nn, tgnn.conv, tgnn.inits = tgnn, tgconv, tginits
for m in (tg, tgnn, tgconv, tginits):
    sys.modules[m.__name__] = m
for name in ("func_helper", "utils", "ecc"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["func_helper"].np = np

import spconv  # noqa: E402  (this repository's package)
assert os.path.realpath(spconv.__file__).startswith(os.path.realpath(ROOT)), spconv.__file__
import backbone_3D_WSIS as ref_backbone  # noqa: E402  (the reference's file)
assert os.path.realpath(ref_backbone.__file__).startswith(os.path.realpath(REF)), ref_backbone.__file__
import sparse_unet3d as ref_unet  # noqa: E402
from wsis_b200 import pipeline  # noqa: E402

cfg = types.SimpleNamespace(**pipeline.DEFAULT_MODEL_CFG)
torch.manual_seed(123)
ref_net = ref_backbone.Network(cfg)
mirror = pipeline.build_network(seed=123, device="cpu")
a, b = ref_net.state_dict(), mirror.state_dict()
same_keys = list(a.keys()) == list(b.keys())
diff = [k for k in a if k in b and not torch.equal(a[k], b[k])]
uses_dropin = all(type(m).__module__.startswith("spconv") for m in ref_net.modules()
                  if type(m).__name__ in ("SubMConv3d", "SparseConv3d", "SparseInverseConv3d", "SparseSequential"))
n_sp = sum(1 for m in ref_net.modules() if isinstance(m, spconv.conv.SparseConvolution))
print(json.dumps({"same_keys": same_keys, "n_keys": len(a), "different_tensors": diff[:5], "uses_dropin": uses_dropin,
                  "sparse_conv_modules": n_sp, "ublock_is_reference": ref_unet.UBlock is type(ref_net.unet)}))
