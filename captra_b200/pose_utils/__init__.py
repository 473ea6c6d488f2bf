"""Device-side mirror of the reference's pose_utils/{procrustes,pose_fit}.py (same names/args)."""
from . import procrustes, pose_fit  # noqa: F401
