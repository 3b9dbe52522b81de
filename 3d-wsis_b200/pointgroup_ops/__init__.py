"""Drop-in for the external `pointgroup_ops` extension (dvlab-research/PointGroup lib/pointgroup_ops; linked
from the reference README.md:37-41, source not in the reference tree).  Contract taken from the call sites:

    voxelization_idx(coords int64[N,4] CPU, batchsize, mode=4) -> (voxel_locs int64[M,4], p2v_map int32[N],
                                                                    v2p_map int32[M,1+maxActive])
                                                        modules/datasets/scannetv2_dataset.py:449,528
    voxelization(feats f32[N,C] cuda, v2p_map int32 cuda, mode=4) -> f32[M,C]   (differentiable)
                                                        train_scannetv2.py:189, test_scannetv2.py:182
"""
import torch
from torch.autograd import Function

from wsis_b200 import ops as W


def voxelization_idx(coords, batchsize, mode=4):
    """mode 4 = mean (config/ScanNet_v2_3D_WSIS.yaml:7); the maps themselves do not depend on the mode."""
    assert coords.dtype == torch.long and coords.dim() == 2 and coords.shape[1] == 4
    return W.voxelization_idx(coords, batchsize, mode)


class Voxelization(Function):
    @staticmethod
    def forward(ctx, feats, map_rule, mode=4):
        assert mode == 4, "only mode 4 (mean) is used by 3D-WSIS"
        assert map_rule.is_contiguous() and feats.is_contiguous()
        ctx.n_points = feats.shape[0]
        ctx.save_for_backward(map_rule)
        return W.voxelization_fwd(feats, map_rule)

    @staticmethod
    def backward(ctx, d_output_feats):
        (map_rule,) = ctx.saved_tensors
        return W.voxelization_bwd(d_output_feats.contiguous(), map_rule, ctx.n_points), None, None


voxelization = Voxelization.apply


def voxelization_backward(d_output_feats, map_rule, n_points, mode=4):
    """Raw backward op (the upstream extension exports it as a function too)."""
    assert mode == 4
    return W.voxelization_bwd(d_output_feats, map_rule, n_points)
