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


def install_dropin(pose=True, mirror_pointnet_lib=None):
    """Make the reference's imports resolve to this package, so its own Python runs unchanged on the B200 kernels:

      `import pointnet2_cuda as pointnet2`          (network/models/pointnet_lib/pointnet2_utils.py:7)
            -> captra_b200.pointnet2_cuda: the reference's OWN pointnet2_utils.py (autograd Functions, QueryAndGroup ...)
               then runs on the drop-in launchers;
      `from pose_utils.procrustes import ...`, `from pose_utils.pose_fit import part_fit_st_no_ransac`
            (networks.py:15-16) and `from procrustes import transform_pts_mask` (pose_fit.py:6)
            -> captra_b200.pose_utils.{procrustes, pose_fit} (pose=True): the torch.svd-on-CPU call sites
               (procrustes.py:27-30,170-174) become device kernels, part_fit_st_no_ransac one fused launch.

      `from pointnet_lib import pointnet2_utils`    (network/models/pointnet_utils.py:10)
            -> the reference's own file when its tree is on sys.path (mirror_pointnet_lib=None / False); this package's
               mirror of it (captra_b200.pointnet_lib) when it is not, or when mirror_pointnet_lib=True.

    Call it BEFORE importing the reference's network.models.* modules (and with the reference tree on sys.path).  The
    rest of the reference's pose_utils package (part_dof_utils, rotations, metrics) stays the reference's own.
    For the fully fused path use the mirrored modules (captra_b200.networks / backbones / pointnet_utils) instead."""
    import importlib.util
    from . import pointnet2_cuda
    sys.modules["pointnet2_cuda"] = pointnet2_cuda
    if mirror_pointnet_lib is None:
        try:
            mirror_pointnet_lib = importlib.util.find_spec("pointnet_lib") is None
        except (ImportError, ValueError):
            mirror_pointnet_lib = True
    if mirror_pointnet_lib:
        from . import pointnet_lib
        from .pointnet_lib import pointnet2_utils
        sys.modules["pointnet_lib"] = pointnet_lib
        sys.modules["pointnet_lib.pointnet2_utils"] = pointnet2_utils
    if pose:
        from .pose_utils import pose_fit, procrustes
        sys.modules["pose_utils.procrustes"] = procrustes
        sys.modules["pose_utils.pose_fit"] = pose_fit
        sys.modules["procrustes"] = procrustes
        sys.modules["pose_fit"] = pose_fit
    return pointnet2_cuda
