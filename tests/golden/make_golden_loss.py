#!/usr/bin/env python
"""Golden vectors for the training loss: the reference's OWN MultiTaskLoss (modules/model/losses_3D_WSIS.py:13-230),
imported unmodified in this container and run on the CPU (its hard-coded `self.device = 'cuda'` attribute is set to
'cpu' on the instance; `import pointgroup_ops` resolves to this repository's drop-in), on a synthetic 2-scene batch
with random-but-fixed network outputs.  Saves inputs, every loss term and the gradient of the total loss w.r.t. every
network output.

    python tests/golden/make_golden_loss.py        # writes tests/golden/loss_batch2.npz
"""
import logging
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200"), "/root/reference/modules/model"):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_inputs(seed=0, n_points=2500):
    from wsis_b200 import synthetic
    rng = np.random.default_rng(seed)
    b = synthetic.collate([synthetic.make_scene(2000 + i, n_points=n_points) for i in range(2)], with_labels=True)
    N, S = b["locs"].shape[0], b["num_superpoints"]
    out = {k: b[k].numpy() for k in ("semantic_labels", "instance_labels", "superpoint_semantic_labels",
                                     "superpoint_instance_labels", "superpoint_offset_vector",
                                     "superpoint_instance_voxel_num", "superpoint_instance_size", "sp_batch_offsets")}
    out["semantic_scores"] = rng.normal(0, 2, (N, 20)).astype(np.float32)
    out["sp_semantic_scores"] = rng.normal(0, 2, (S, 20)).astype(np.float32)
    out["pred_sp_offset_vectors"] = rng.normal(0, 0.5, (S, 3)).astype(np.float32)
    out["pred_sp_occupancy"] = rng.normal(5, 2, S).astype(np.float32)
    out["pred_sp_ins_size"] = rng.normal(1, 1, S).astype(np.float32)
    # instance-clustered discriminative features so that every hinge of the loss is active somewhere
    inst = out["superpoint_instance_labels"]
    base = rng.normal(0, 1.0, (int(inst.max()) + 2, 7))
    out["sp_discriminative_feats"] = (base[np.maximum(inst, -1) + 1] + rng.normal(0, 0.3, (S, 7))).astype(np.float32)
    return out


PRED = ("semantic_scores", "sp_semantic_scores", "pred_sp_offset_vectors", "pred_sp_occupancy", "pred_sp_ins_size",
        "sp_discriminative_feats")


def main():
    import losses_3D_WSIS as ref                                      # the reference's file
    assert ref.__file__.startswith("/root/reference"), ref.__file__
    inp = make_inputs()
    logger = logging.getLogger("golden")
    out = dict(inp)
    for tag, epoch in (("early", 1), ("joint", 121)):
        crit = ref.MultiTaskLoss(logger, SimpleNamespace(ignore_label=-100, supervise_instance_size=True,
                                                         joint_training_epoch=120, semantic_dice=True,
                                                         supervise_sp_offset=True), SimpleNamespace(classes=20))
        crit.device = 'cpu'
        t = {k: torch.from_numpy(inp[k]).clone().requires_grad_(k in PRED) for k in inp}
        li = {'point_labels': (t["semantic_labels"], t["instance_labels"]), "semantic_scores": t["semantic_scores"],
              'superpoint_labels': (t["superpoint_semantic_labels"], t["superpoint_instance_labels"]),
              'sp_semantic': t["sp_semantic_scores"],
              'sp_offset_vector': (t["pred_sp_offset_vectors"], t["superpoint_offset_vector"]),
              'sp_occupancy': (t["pred_sp_occupancy"], t["superpoint_instance_voxel_num"]),
              'sp_instance_size': (t["pred_sp_ins_size"], t["superpoint_instance_size"]),
              'sp_discriminative_features': (t["sp_discriminative_feats"], t["sp_batch_offsets"])}
        loss, parts = crit(li, epoch)
        loss.backward()
        out["loss_" + tag] = np.float64(loss.item())
        for k, v in parts.items():
            out["%s_%s" % (k, tag)] = np.float64(v[0].item())
        for k in PRED:
            if t[k].grad is not None:
                out["grad_%s_%s" % (k, tag)] = t[k].grad.numpy()
        print(tag, "loss", loss.item(), {k: round(float(v[0]), 5) for k, v in parts.items()})
    np.savez_compressed(os.path.join(HERE, "loss_batch2.npz"), **out)


if __name__ == "__main__":
    main()
