"""numpy float64 restatement of pose_utils/procrustes.py and pose_utils/pose_fit.py.

TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.c header).  Each function follows the cited
reference lines operation by operation, in float64, with numpy.linalg.svd standing in for
torch.svd (LAPACK gesdd in both; U, V are sign-ambiguous, only R = U diag(1,..,det) V^T is
compared -- SURVEY.md section 8c).  Pinned against tests/golden/procrustes.npz, which was
produced by running the reference's own functions (tests/golden/make_golden.py).
"""
import numpy as np

EPS = 1e-6  # procrustes.py:5


def _T(a):
    return np.swapaxes(a, -1, -2)


def rotate_pts_batch(source, target):
    """procrustes.py:25-56."""
    M = _T(target) @ source
    U, _, Vh = np.linalg.svd(M)
    d = np.linalg.det(U @ Vh)
    mid = np.zeros_like(U)
    mid[..., 0, 0] = 1.0
    mid[..., 1, 1] = 1.0
    mid[..., 2, 2] = d
    return U @ mid @ Vh


def rotate_pts_2d_batch(source, target):
    """procrustes.py:167-204 (validity check in float32 like the reference)."""
    M = _T(target) @ source
    U, _, Vh = np.linalg.svd(M)
    d = np.linalg.det(U @ Vh)
    mid = np.zeros_like(U)
    mid[..., 0, 0] = 1.0
    mid[..., 1, 1] = d
    R = U @ mid @ Vh
    res = np.abs((_T(R) @ R).astype(np.float32) - np.eye(2, dtype=np.float32)).mean(axis=(-1, -2))
    ok = (res < 1e-5)[..., None, None]
    return np.where(ok, R, np.eye(2))


def rot_around_yaxis_to_3d(r2):
    """procrustes.py:69-75."""
    out = np.zeros(r2.shape[:-2] + (3, 3))
    out[..., 0, 0], out[..., 0, 2] = r2[..., 0, 0], r2[..., 0, 1]
    out[..., 1, 1] = 1.0
    out[..., 2, 0], out[..., 2, 2] = r2[..., 1, 0], r2[..., 1, 1]
    return out


def scale_pts_batch(source, target):
    """procrustes.py:59-62."""
    return (source * target).sum((-1, -2)) / ((source * source).sum((-1, -2)) + EPS)


def translate_pts_batch(source, target):
    """procrustes.py:65-66."""
    return (target - source).mean(-1, keepdims=True)


def scale_pts_mask(source, target, w):
    """procrustes.py:117-120."""
    return (source * target * w).sum((-1, -2)) / ((source * source * w).sum((-1, -2)) + EPS)


def translate_pts_mask(source, target, w):
    """procrustes.py:123-129."""
    w = _T(w)
    wn = w / np.maximum(w.sum(-1, keepdims=True), 1.0)
    return ((target - source) * wn).sum(-1, keepdims=True)


def _center(x, mask):
    return (x * mask).sum(-2, keepdims=True) / np.maximum(mask.sum(-2, keepdims=True), 1.0)


def transform_pts_2d_mask(source, target, mask):
    """procrustes.py:213-228."""
    sc = (source - _center(source, mask)) * mask
    tc = (target - _center(target, mask)) * mask
    R = rotate_pts_2d_batch(sc * mask, tc * mask)  # rotate_pts_2d_mask :207-210
    t = translate_pts_mask(R @ _T(source), _T(target), mask)
    return R, t


def transform_pts_2d_batch(source, target):
    """procrustes.py:231-242."""
    sc = source - source.mean(-2, keepdims=True)
    tc = target - target.mean(-2, keepdims=True)
    R = rotate_pts_2d_batch(sc, tc)
    return R, translate_pts_batch(R @ _T(source), _T(target))


def transform_pts_batch(source, target, given_scale=None, rotation=None, sym=False):
    """procrustes.py:78-107."""
    source, target = np.asarray(source, np.float64), np.asarray(target, np.float64)
    sc = source - source.mean(-2, keepdims=True)
    tc = target - target.mean(-2, keepdims=True)
    if rotation is None:
        rotation = rotate_pts_batch(sc, tc)
    if sym:
        canon = target @ rotation
        r2, _ = transform_pts_2d_batch(source[..., [0, 2]], canon[..., [0, 2]])
        rotation = rotation @ rot_around_yaxis_to_3d(r2)
    scale = given_scale if given_scale is not None else scale_pts_batch(sc @ _T(rotation), tc)
    t = translate_pts_batch(scale[..., None, None] * (rotation @ _T(source)), _T(target))
    return rotation, scale, t


def transform_pts_mask(source, target, mask, weights, given_scale=None, rotation=None, sym=False):
    """procrustes.py:132-164."""
    source, target = np.asarray(source, np.float64), np.asarray(target, np.float64)
    mask, weights = np.asarray(mask, np.float64), np.asarray(weights, np.float64)
    sc = (source - _center(source, mask)) * mask
    tc = (target - _center(target, mask)) * mask
    if rotation is None:
        w = np.sqrt(weights + EPS)  # rotate_pts_mask :110-114
        rotation = rotate_pts_batch(sc * w, tc * w)
    rotation = np.asarray(rotation, np.float64)
    if sym:
        canon = target @ rotation
        r2, _ = transform_pts_2d_mask(source[..., [0, 2]], canon[..., [0, 2]], weights)
        rotation = rotation @ rot_around_yaxis_to_3d(r2)
    scale = given_scale if given_scale is not None else scale_pts_mask(sc @ _T(rotation), tc, weights)
    t = translate_pts_mask(scale[..., None, None] * (rotation @ _T(source)), _T(target), weights)
    return rotation, scale, t


def part_fit_st_no_ransac(labels, source, target, rotation, cfg, given_scale=None):
    """pose_fit.py:38-53 (+ filter_model_valid :26-35)."""
    P = cfg["num_parts"]
    eye = np.concatenate([np.eye(P), np.zeros((2, P))], 0)
    mask = _T(eye[labels])                       # [B,P,N]
    valid = mask.sum(-1) > 3
    _, scale, t = transform_pts_mask(source, target, mask[..., None], mask[..., None],
                                     given_scale=given_scale, rotation=rotation, sym=cfg["sym"])
    valid &= np.isfinite(scale) & np.isfinite(t.sum((-1, -2)))
    if rotation is not None:
        valid &= np.isfinite(np.asarray(rotation).sum((-1, -2)))
    return {"rotation": rotation, "scale": scale, "translation": t}, valid
