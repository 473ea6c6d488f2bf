"""Mirror of pose_utils/procrustes.py -- same function names, argument shapes and return values
(reference line numbers cited per function) with the two CPU round trips removed: the 3x3 and
2x2 `torch.svd` calls that the reference runs after `M.cpu()` (procrustes.py:27-30,170-174) are
closed-form device kernels (csrc/pose_fit.cu: captra_procrustes_rot3 / captra_procrustes_rot2).

Everything else is elementwise/reduction torch code that already runs on the device in the
reference; it is restated here, not imported, so that the package is self-contained.  The
tracker's hot call, part_fit_st_no_ransac, bypasses this composition and uses one fused kernel
(see pose_fit.py).

Autograd: gradients flow through scale/translation/centring exactly as in the reference
(networks.py:88-100 differentiates scale_pts_mask / translate_pts_mask); the rotation kernels are
non-differentiable -- the reference detaches the 2-D one itself (procrustes.py:169,195) and its
live paths never differentiate the 3-D one (rotation is always given, networks.py:227).
"""
import torch

from .. import _lib

EPS = 1e-6


def _rot_kernel(M, dim):
    """M [..., dim, dim] (any float dtype/layout on CUDA) -> R, same shape, via the C ABI."""
    if not M.is_cuda:
        raise _lib.CaptraError("procrustes: tensors must be on a CUDA device (no CPU path)")
    Mc = M.detach().to(torch.float32).contiguous()
    R = torch.empty_like(Mc)
    count = Mc.numel() // (dim * dim)
    fn = _lib.load().captra_procrustes_rot3 if dim == 3 else _lib.load().captra_procrustes_rot2
    _lib.check(fn(count, Mc.data_ptr(), R.data_ptr(), _lib.stream_ptr(M.device)), "procrustes_rot%d" % dim)
    return R.to(M.dtype)


def rotate_pts_batch(source, target):
    """procrustes.py:25-56.  src, tgt [..., N, 3] -> R [..., 3, 3] = U diag(1,1,det(UV^T)) V^T of
    M = tgt^T src."""
    M = torch.matmul(target.transpose(-1, -2), source)
    return _rot_kernel(M, 3)


def scale_pts_batch(source, target):
    """procrustes.py:59-62."""
    return torch.sum(source * target, dim=(-1, -2)) / (torch.sum(source * source, dim=(-1, -2)) + EPS)


def translate_pts_batch(source, target):
    """procrustes.py:65-66.  [..., 3, N] -> [..., 3, 1]."""
    return torch.mean(target - source, dim=-1, keepdim=True)


def rot_around_yaxis_to_3d(rot_2d):
    """procrustes.py:69-75.  [[xx,xz],[zx,zz]] -> rotation about y."""
    xx, xz, zx, zz = rot_2d[..., 0, 0], rot_2d[..., 0, 1], rot_2d[..., 1, 0], rot_2d[..., 1, 1]
    one, zero = torch.ones_like(xx), torch.zeros_like(xx)
    rows = torch.stack([xx, zero, xz, zero, one, zero, zx, zero, zz], dim=-1)
    return rows.reshape(rows.shape[:-1] + (3, 3))


def transform_pts_batch(source, target, given_scale=None, rotation=None, sym=False):
    """procrustes.py:78-107.  src, tgt [B,P,H,N,3] -> (R [...,3,3], s [...], t [...,3,1])."""
    source_centered = source - torch.mean(source, -2, keepdim=True)
    target_centered = target - torch.mean(target, -2, keepdim=True)
    if rotation is None:
        rotation = rotate_pts_batch(source_centered, target_centered)
    if sym:
        canon_target = torch.matmul(target, rotation)
        rot_2d, _ = transform_pts_2d_batch(source[..., [0, 2]], canon_target[..., [0, 2]])
        rotation = torch.matmul(rotation, rot_around_yaxis_to_3d(rot_2d))
    if given_scale is not None:
        scale = given_scale
    else:
        scale = scale_pts_batch(torch.matmul(source_centered, rotation.transpose(-1, -2)), target_centered)
    translation = translate_pts_batch(
        scale.reshape(scale.shape + (1, 1)) * torch.matmul(rotation, source.transpose(-1, -2)),
        target.transpose(-1, -2))
    return rotation, scale, translation


def rotate_pts_mask(source, target, w):
    """procrustes.py:110-114 (inputs already centred and masked)."""
    w = torch.sqrt(w + EPS)
    return rotate_pts_batch(source * w, target * w)


def scale_pts_mask(source, target, w):
    """procrustes.py:117-120."""
    return (torch.sum(source * target * w, dim=(-1, -2)) /
            (torch.sum(source * source * w, dim=(-1, -2)) + EPS))


def translate_pts_mask(source, target, w):
    """procrustes.py:123-129.  source/target [..., 3, N], w [..., N, 1] -> [..., 3, 1]."""
    w_shape = list(w.shape)
    w_shape[-2], w_shape[-1] = w_shape[-1], w_shape[-2]
    w = w.reshape(w_shape)
    w_normalized = w / torch.clamp(torch.sum(w, dim=-1, keepdim=True), min=1.0)
    return torch.sum((target - source) * w_normalized, dim=-1, keepdim=True)


def _masked_center(x, mask):
    return torch.sum(x * mask, dim=-2, keepdim=True) / torch.clamp(torch.sum(mask, dim=-2, keepdim=True), min=1.0)


def transform_pts_mask(source, target, mask, weights, given_scale=None, rotation=None, sym=False):
    """procrustes.py:132-164.  Masked/weighted Procrustes: (R, s, t)."""
    source_centered = (source - _masked_center(source, mask)) * mask
    target_centered = (target - _masked_center(target, mask)) * mask
    if rotation is None:
        rotation = rotate_pts_mask(source_centered, target_centered, weights)
    if sym:
        canon_target = torch.matmul(target, rotation)
        rot_2d, _ = transform_pts_2d_mask(source[..., [0, 2]], canon_target[..., [0, 2]], weights)
        rotation = torch.matmul(rotation, rot_around_yaxis_to_3d(rot_2d))
    if given_scale is not None:
        scale = given_scale
    else:
        scale = scale_pts_mask(torch.matmul(source_centered, rotation.transpose(-1, -2)), target_centered, weights)
    translation = translate_pts_mask(
        scale.reshape(scale.shape + (1, 1)) * torch.matmul(rotation, source.transpose(-1, -2)),
        target.transpose(-1, -2), weights)
    return rotation, scale, translation


def rotate_pts_2d_batch(source, target):
    """procrustes.py:167-204.  src, tgt [..., N, 2] -> detached R [..., 2, 2] (identity where the
    result fails the reference's orthogonality check)."""
    M = torch.matmul(target.transpose(-1, -2), source).detach()
    return _rot_kernel(M, 2)


def rotate_pts_2d_mask(source, target, w):
    """procrustes.py:207-210."""
    return rotate_pts_2d_batch(source * w, target * w)


def transform_pts_2d_mask(source, target, mask):
    """procrustes.py:213-228.  src, tgt [B,P,N,2], mask [B,P,N,1] -> (R [B,P,2,2], t [B,P,2,1])."""
    source_centered = (source - _masked_center(source, mask)) * mask
    target_centered = (target - _masked_center(target, mask)) * mask
    rotation = rotate_pts_2d_mask(source_centered, target_centered, mask)
    translation = translate_pts_mask(torch.matmul(rotation, source.transpose(-1, -2)),
                                     target.transpose(-1, -2), mask)
    return rotation, translation


def transform_pts_2d_batch(source, target):
    """procrustes.py:231-242."""
    source_centered = source - torch.mean(source, -2, keepdim=True)
    target_centered = target - torch.mean(target, -2, keepdim=True)
    rotation = rotate_pts_2d_batch(source_centered, target_centered)
    translation = translate_pts_batch(torch.matmul(rotation, source.transpose(-1, -2)), target.transpose(-1, -2))
    return rotation, translation
