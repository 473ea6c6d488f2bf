"""captra_b200 -- B200-native (sm_100a) implementation of CAPTRA's per-frame point-cloud hot path.

Layout (only what the path needs):
  csrc/                      hand-written CUDA kernels + the C ABI (include/captra_ops.h)
  _lib.py                    ctypes loader for libcaptra_ops.so (no fallback: raises if missing)
  pointnet2_cuda.py          drop-in for the reference's pybind module of the same name
  pointnet_lib/              mirror of network/models/pointnet_lib/pointnet2_utils.py
  pointnet_utils.py          mirror of network/models/pointnet_utils.py (SA-MSG / FP / group-all)
  backbones.py               mirror of network/models/backbones.py (PointNet2Msg)
  pose_utils/                mirror of pose_utils/{procrustes,pose_fit}.py on the device
"""
import sys

__version__ = "0.1.0"


def install_dropin():
    """Make `import pointnet2_cuda`, `from pointnet_lib import pointnet2_utils` and
    `from pose_utils import procrustes, pose_fit` resolve to this package, so the reference's
    network/models/pointnet_utils.py:10, pointnet_lib/pointnet2_utils.py:7 and networks.py:15-16
    run unchanged on the B200 kernels."""
    from . import pointnet2_cuda, pointnet_lib
    from .pointnet_lib import pointnet2_utils
    sys.modules["pointnet2_cuda"] = pointnet2_cuda
    sys.modules["pointnet_lib"] = pointnet_lib
    sys.modules["pointnet_lib.pointnet2_utils"] = pointnet2_utils
    return pointnet2_cuda
