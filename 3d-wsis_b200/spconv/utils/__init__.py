"""The reference's spconv.utils holds KITTI detection helpers (NMS, rotated IoU, points_to_voxel) that 3D-WSIS
never calls (out of scope); the module exists so `from spconv import utils` keeps working."""
