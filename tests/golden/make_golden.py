"""Generates tests/golden/*.npz by executing the UNMODIFIED reference code in the build container.

    python tests/golden/make_golden.py            (needs /root/reference and oracle/_ref/libspconv_ref.so)

What runs is the reference itself: modules/model/backbone_3D_WSIS.py `Network` (+ sparse_unet3d.py, graphnet.py,
spg_modules.py) on top of the reference's own spconv Python package and its CPU kernels compiled from source
(oracle/Makefile).  What is NOT in the reference tree or not installed here is stubbed with the restatements
the SURVEY names (§8c) -- these pieces are "parity unpinned":
    torch_scatter.scatter      -> index_add_/amax formulation
    torch_geometric NNConv     -> MessagePassing stub (flow source_to_target -- spg_modules.py:68 never forwards `flow` -- aggr mean)
    pointgroup_ops             -> oracle.voxelization_idx / voxelization
    ecc.GraphConvInfo (igraph) -> tensor-only stand-in with the same get_buffers()/get_pyg_buffers()
    func_helper, utils, ecc    -> empty modules (imported by the reference, unused on this path)
The reference tree does not exist on the GPU box, so the outputs are committed as fixtures next to this script.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("WSIS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def install_reference_imports():
    """Makes `import spconv`, `import backbone_3D_WSIS` resolve to the reference's files."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libspconv_ref.so")
    assert os.path.exists(ref_so), "run `make -C oracle ref` first"
    shadow = tempfile.mkdtemp(prefix="wsis_refpkg_")
    pkg = os.path.join(shadow, "spconv")
    os.makedirs(pkg)
    src = os.path.join(REF, "modules/lib/spconv/spconv")
    for f in ("__init__.py", "conv.py", "functional.py", "modules.py", "ops.py", "pool.py", "test_utils.py"):
        os.symlink(os.path.join(src, f), os.path.join(pkg, f))
    os.symlink(ref_so, os.path.join(pkg, "libspconv.so"))
    sys.modules["spconv.utils"] = types.ModuleType("spconv.utils")  # real one needs the boost pybind lib
    sys.path.insert(0, shadow)
    sys.path.insert(0, os.path.join(REF, "modules/model"))

    from oracle import oracle as orc

    # ---- torch_scatter restatement -------------------------------------------------------------------
    ts = types.ModuleType("torch_scatter")

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
        assert dim == 0
        S = int(index.max()) + 1 if dim_size is None else dim_size
        shape = (S,) + tuple(src.shape[1:])
        if reduce == "max":
            idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
            return torch.zeros(shape, dtype=src.dtype).scatter_reduce(0, idx, src, "amax", include_self=False)
        res = torch.zeros(shape, dtype=src.dtype).index_add_(0, index, src)
        if reduce == "mean":
            cnt = torch.zeros(S, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
            res = res / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
        return res

    ts.scatter = scatter
    ts.scatter_add = lambda src, index, dim=0, **kw: scatter(src, index, dim, reduce="sum", **kw)
    ts.scatter_mean = lambda src, index, dim=0, **kw: scatter(src, index, dim, reduce="mean", **kw)
    ts.scatter_max = ts.scatter_min = None
    sys.modules["torch_scatter"] = ts

    # ---- torch_geometric MessagePassing stub ------------------------------------------------------------
    tg, tgnn, tgconv, tginits = (types.ModuleType(n) for n in
                                 ("torch_geometric", "torch_geometric.nn", "torch_geometric.nn.conv",
                                  "torch_geometric.nn.inits"))

    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add", flow="source_to_target", **kw):
            super().__init__()
            self.aggr, self.flow = aggr, flow

        def propagate(self, edge_index, x, weights):
            i, j = (0, 1) if self.flow == "target_to_source" else (1, 0)
            msg = self.message(edge_index[i], x[edge_index[j]], x.size(0), weights)
            out = scatter(msg, edge_index[i], 0, dim_size=x.size(0), reduce={"add": "sum"}.get(self.aggr, self.aggr))
            return self.update(out, x)

    tgconv.MessagePassing = MessagePassing
    tginits.uniform = lambda size, tensor: None if tensor is None else tensor.data.uniform_(-1 / size ** 0.5, 1 / size ** 0.5)
    tg.nn, tgnn.conv, tgnn.inits = tgnn, tgconv, tginits
    for m in (tg, tgnn, tgconv, tginits):
        sys.modules[m.__name__] = m

    # ---- pointgroup_ops restatement (oracle) ------------------------------------------------------------
    pg = types.ModuleType("pointgroup_ops")

    def voxelization_idx(coords, batchsize, mode=4):
        return tuple(torch.from_numpy(a) for a in orc.voxelization_idx(coords.numpy(), batchsize, mode))

    pg.voxelization_idx = voxelization_idx
    pg.voxelization = lambda feats, v2p, mode=4: torch.from_numpy(orc.voxelization(feats.numpy(), v2p.numpy(), mode))
    sys.modules["pointgroup_ops"] = pg

    for name in ("func_helper", "utils", "ecc"):
        sys.modules[name] = types.ModuleType(name)
    # backbone_3D_WSIS.py uses `np` without importing it: it arrives through `from func_helper import *`
    sys.modules["func_helper"].np = np


class RefGraphInfo(object):
    def __init__(self, edge_index, edgefeats):
        self._edge_indexes, self._edgefeats = edge_index, edgefeats

    def cuda(self):
        pass

    def get_buffers(self):
        return None, None, None, None, self._edgefeats

    def get_pyg_buffers(self):
        return self._edge_indexes


def reference_forward(net, batch):
    """train_scannetv2.py:149-198 on CPU tensors, with the reference Network."""
    import pointgroup_ops
    import spconv
    from torch_scatter import scatter
    voxel_locs, p2v, v2p = pointgroup_ops.voxelization_idx(batch["locs"], batch["batch_size"], 4)
    centers = scatter(batch["locs_float"], batch["superpoint"], dim=0, reduce="mean")
    feats = torch.cat((batch["feats"], batch["locs_float"]), 1)
    voxel_feats = pointgroup_ops.voxelization(feats, v2p, 4)
    inp = spconv.SparseConvTensor(voxel_feats, voxel_locs.int(), batch["spatial_shape"], batch["batch_size"])
    extra = {"superpoint": batch["superpoint"], "GIs": [RefGraphInfo(batch["ecc_edge_index"], batch["ecc_edgefeats"])],
             "edge_u_list": batch["edge_u_list"], "edge_v_list": batch["edge_v_list"],
             "superpoint_cenetr_xyz": centers}
    # the reference hard-codes cuda=True in `self.ecc.set_info(GIs, cuda=True)`; RefGraphInfo.cuda() is a no-op
    unet_out = {}
    h = net.output_layer.register_forward_hook(lambda m, i, o: unet_out.__setitem__("f", o.features.detach().clone()))
    ret = net(inp, p2v, extra)
    h.remove()
    rb = {k: v for k, v in inp.indice_dict.items()}
    return ret, dict(voxel_locs=voxel_locs, p2v=p2v, v2p=v2p, voxel_feats=voxel_feats, centers=centers,
                     unet_out=unet_out["f"], rulebooks=rb)


def main():
    install_reference_imports()
    # our synthetic generator is imported by file path so that OUR spconv package is never imported here
    import importlib.util
    spec = importlib.util.spec_from_file_location("wsis_synth", os.path.join(ROOT, "3d-wsis_b200/wsis_b200/synthetic.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)

    from types import SimpleNamespace
    import backbone_3D_WSIS as ref_model  # the reference's file
    torch.set_num_threads(8)

    cfg = SimpleNamespace(input_channel=3, use_coords=True, blocks=5, block_reps=2, media=32, classes=20, fix_module="[]")
    torch.manual_seed(123)
    net = ref_model.Network(cfg).eval()
    # non-trivial BatchNorm statistics so the folded-BN fusion is really exercised
    g = torch.Generator().manual_seed(7)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
            if m.affine:
                m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)

    for name, nscene, kw in (("scene_b1", 1, dict(n_points=6000, room=(1.0, 0.8, 0.6), n_boxes=2, sp_cell=0.1, box_scale=0.25)),
                             ("scene_b2", 2, dict(n_points=3000, room=(0.8, 0.6, 0.5), n_boxes=1, sp_cell=0.1, box_scale=0.2))):
        scenes = [synth.make_scene(9000 + i, **kw) for i in range(nscene)]
        batch = synth.collate(scenes)
        with torch.no_grad():
            ret, aux = reference_forward(net, batch)
        out = {"ret_" + k: v.numpy() for k, v in ret.items()}
        out.update(voxel_locs=aux["voxel_locs"].numpy(), p2v=aux["p2v"].numpy(), v2p=aux["v2p"].numpy(),
                   voxel_feats=aux["voxel_feats"].numpy(), centers=aux["centers"].numpy(),
                   unet_out=aux["unet_out"].numpy())
        for key in ("subm1", "spconv1", "subm2", "spconv2"):
            outids, _, pairs, num, _ = aux["rulebooks"][key]
            out["rb_%s_outids" % key] = outids.numpy()
            out["rb_%s_num" % key] = num.numpy()
            if key in ("subm1", "spconv1"):
                out["rb_%s_pairs" % key] = pairs.numpy()
        out["param_checksum"] = np.array([float(p.double().abs().sum()) for p in net.parameters()])
        out["param_names"] = np.array([n for n, _ in net.named_parameters()])
        out["scene_kw"] = np.array(repr((nscene, kw)))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "N", batch["locs"].shape[0], "M", aux["voxel_locs"].shape[0], "S", batch["num_superpoints"],
              "E", batch["edge_u_list"].shape[0], {k: tuple(v.shape) for k, v in ret.items()})
    sd_keys = np.array(list(net.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "state_dict_keys.npz"), keys=sd_keys,
                        shapes=np.array([repr(tuple(v.shape)) for v in net.state_dict().values()]))


if __name__ == "__main__":
    main()
