"""B200-native host runtime for the 3D-WSIS scene-level hot path.

Layout of the directory this package lives in (`3d-wsis_b200/`, put it on sys.path the way the reference puts
`modules/lib` there):
    spconv/          drop-in for the reference's spconv Python package  (modules/lib/spconv/spconv/*.py)
    pointgroup_ops/  drop-in for the external pointgroup_ops extension   (README.md:37-41)
    wsis_b200/       this package: C-ABI loader, tensor-level ops, the host mirror of the model, synthetic scenes
    csrc/            the sm_100a CUDA sources of libwsis_b200.so (include/wsis_b200.h)
"""
from . import ops  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
