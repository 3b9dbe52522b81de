"""Multi-GPU plumbing: one process per GPU under torchrun, `torch.distributed` over NCCL (gloo in CPU tests).

The hot path shards by SCENE: every stage is per-scene independent (the batch index is the leading
voxel coordinate and rulebooks never cross it, indice.cu.h:164-166), so inference places whole scenes on ranks with
no data-path collective; the only collective is the per-scene result gather at the end of an epoch (host objects,
like the reference's utils/comm.py:145-225 helpers).  Training is data parallel: the reference wraps the model in
DistributedDataParallel (train_scannetv2.py:734-738) but never creates the process group; here the gradients of
the (small: 11 M parameters = 44 MB fp32) model are flattened into ONE bucket and all-reduced once per step --
at NVSwitch bandwidth that is latency bound, so one launch beats DDP's 25 MB buckets + unused-parameter search.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None, device_index=None):
    """Creates the default process group from the torchrun environment; returns (rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_scenes(n_scenes, rank=None, world_size=None):
    """Scene ids owned by `rank`: round-robin, so every rank gets ceil/floor(n/world) scenes and consecutive
    (similar-sized) scans spread across GPUs.  The union over ranks is exactly range(n_scenes)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, n_scenes, world_size))


def gather_scene_results(local, n_scenes):
    """local: {scene_id: picklable result} of this rank -> list of length n_scenes on every rank (the per-scene
    pseudo labels / predictions that the reference writes back into its dataset objects,
    train_scannetv2.py:575-581).  No-op on one rank."""
    rank, world_size = world()
    if world_size == 1:
        parts = [local]
    else:
        parts = [None] * world_size
        dist.all_gather_object(parts, local)
    out = [None] * n_scenes
    for part in parts:
        for sid, res in part.items():
            assert out[sid] is None, "scene %d produced by two ranks" % sid
            out[sid] = res
    missing = [i for i, r in enumerate(out) if r is None]
    assert not missing, "scenes %s were not produced by any rank" % missing[:8]
    return out


class GradBucket:
    """One flat fp32 buffer aliasing every parameter's .grad; `allreduce()` averages it over the ranks with a
    single collective.  Parameters that did not receive a gradient this step contribute zeros (the reference
    needs find_unused_parameters=True for the same reason: heads are unused before joint training,
    losses_3D_WSIS.py:68)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=dt, device=dev)
        off = 0
        for p, n in zip(self.params, self.sizes):
            p.grad = self.flat[off:off + n].view_as(p)  # autograd accumulates in place into the bucket
            off += n

    def zero(self):
        self.flat.zero_()
        off = 0
        for p, n in zip(self.params, self.sizes):  # optimizers that set grads to None: re-alias
            view = self.flat[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += n

    def check_aliasing(self):
        """Every parameter's .grad must still be its view of the flat buffer: `optimizer.zero_grad(set_to_none=True)`
        (the torch default) drops the views, autograd then allocates fresh gradients and the all-reduce of the bucket
        would silently synchronise nothing.  Use `bucket.zero()` instead of `optimizer.zero_grad()`."""
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is not None and p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                raise RuntimeError("GradBucket: a parameter's .grad no longer aliases the flat bucket (was "
                                   "optimizer.zero_grad(set_to_none=True) called?); call bucket.zero() instead")
            off += n

    def allreduce(self, async_op=False):
        self.check_aliasing()
        _, world_size = world()
        if world_size == 1:
            return None
        self.flat.div_(world_size)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
