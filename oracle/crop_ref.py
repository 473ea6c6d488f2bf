"""CPU restatement (numpy, float64) of the reference's per-frame crop -- TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.c).
Follows datasets/nocs_data/nocs_utils.py:5-33 (backproject), :36-42 (project), :57-77 (get_corners, bbox_from_corners),
datasets/nocs_data/nocs_data_process.py:92-109 (crop_ball_from_pts), :129-143 (get_proj_corners), :148-164
(crop_ball_from_depth_image) and datasets/data_utils.py:138-158 (farthest_point_sample: CUDA branch -- a random
5 * npoint subset, then FPS from index 0 with the CUDA kernel's semantics, here oracle/cpu_ref).  The reference draws
the subset from numpy's global RNG; here the permutation is an explicit argument.  Pinned against the reference's own
functions by tests/golden/crop.npz (tests/golden/make_golden.py::golden_crop)."""
import numpy as np

from . import cpu_ref


def backproject(depth, intrinsics, mask=None, scale=0.001):
    intrinsics_inv = np.linalg.inv(intrinsics)
    height = depth.shape[0]
    non_zero_mask = depth > 0
    final = np.logical_and(mask, non_zero_mask) if mask is not None else non_zero_mask
    idxs = np.where(final)
    grid = np.array([idxs[1], height - idxs[0]])
    uv = np.concatenate((grid, np.ones([1, grid.shape[1]])), axis=0)
    xyz = np.transpose(intrinsics_inv @ uv)
    z = depth[idxs[0], idxs[1]].astype(np.float32)
    pts = xyz * z[:, np.newaxis] / xyz[:, -1:]
    pts[:, 2] = -pts[:, 2]
    return pts * scale, idxs


def project(pts, intrinsics, scale=1000):
    pts = pts * scale
    pts = -pts / pts[:, -1:]
    pts[:, -1] = -pts[:, -1]
    return np.transpose(intrinsics @ np.transpose(pts))[:, :2]


def get_proj_corners(depth, center, radius, cam_intrinsics):
    radius = max(radius, 0.05)
    cr = np.stack([center - np.ones(3) * radius, center + np.ones(3) * radius])
    aabb = np.zeros((8, 3))
    for i in range(8):
        x, y, z = (i % 4) // 2, i // 4, i % 2
        aabb[i] = (cr[x, 0], cr[y, 1], cr[z, 2])
    height, width = depth.shape
    pc = project(aabb, cam_intrinsics).astype(np.int32)[:, [1, 0]]
    pc[:, 0] = height - pc[:, 0]
    c2 = np.stack([np.min(pc, axis=0), np.max(pc, axis=0)], axis=0)
    c2[0, :] = np.maximum(c2[0, :], 0)
    c2[1, :] = np.minimum(c2[1, :], np.array([height - 1, width - 1]))
    return c2


def farthest_point_sample(xyz, npoint, perm=None):
    """data_utils.py:138-158, CUDA branch."""
    if len(xyz) > 5 * npoint:
        idx = np.asarray(perm)[:5 * npoint]
        sub = np.ascontiguousarray(xyz[idx].astype(np.float32)).reshape(1, -1, 3)
        return idx[cpu_ref.furthest_point_sample(sub, npoint).reshape(-1)]
    return cpu_ref.furthest_point_sample(np.ascontiguousarray(xyz.astype(np.float32)).reshape(1, -1, 3), npoint).reshape(-1)


def crop_ball_from_pts(pts, center, radius, num_points=None, perm=None):
    distance = np.sqrt(np.sum((pts - center) ** 2, axis=-1))
    radius = max(radius, 0.05)
    for _ in range(10):
        idx = np.where(distance <= radius)[0]
        if len(idx) >= 10 or num_points is None:
            break
        radius *= 1.10
    if num_points is not None:
        if len(idx) == 0:
            idx = np.where(distance <= 1e9)[0]
        if len(idx) == 0:
            return idx
        while len(idx) < num_points:
            idx = np.concatenate([idx, idx], axis=0)
        idx = idx[farthest_point_sample(pts[idx], num_points, perm)]
    return idx


def crop_ball_from_depth_image(depth, mask, center, radius, cam_intrinsics, num_points=None, perm=None):
    c2 = get_proj_corners(depth, center, radius, cam_intrinsics)
    corner_mask = np.zeros_like(depth)
    corner_mask[c2[0, 0]: c2[1, 0] + 1, c2[0, 1]: c2[1, 1] + 1] = 1
    raw_pts, raw_idx = backproject(depth, intrinsics=cam_intrinsics, mask=corner_mask)
    raw_mask = mask[raw_idx[0], raw_idx[1]]
    idx = crop_ball_from_pts(raw_pts, center, radius, num_points, perm)
    if len(idx) == 0:
        return crop_ball_from_depth_image(depth, mask, center, radius * 1.2, cam_intrinsics, num_points, perm)
    return raw_pts[idx], raw_mask[idx], idx
