"""Device-side mirror of the reference's per-frame crop, same function names and argument
meaning as datasets/nocs_data/nocs_data_process.py:92-109,129-164 and datasets/nocs_data/nocs_utils.py:5-47:

    get_proj_corners(depth, center, radius, cam_intrinsics)                       host arithmetic on 8 corners
    crop_ball_from_depth_image(depth, mask, center, radius, cam_intrinsics, num_points, device, perm=None)

depth / mask are CUDA tensors [H,W] (depth float32 in millimetres; mask int32) -- the depth window never goes back to
the host; one 12-byte read (the number of selected points) is the only synchronisation.  The reference thins crops of
more than 5 * num_points points with numpy's global RNG (data_utils.py:147); pass `perm` (a permutation of the tiled
list's length, e.g. the one the reference drew) to reproduce it, else torch.randperm on the device is used.
"""
import ctypes

import numpy as np
import torch

from . import _lib, fused_ops

nocs_real_cam_intrinsics = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])   # nocs_data_process.py:20


def project(pts, intrinsics, scale=1000):
    """nocs_utils.py:36-42 (not flipping the y axis)."""
    pts = pts * scale
    pts = -pts / pts[:, -1:]
    pts[:, -1] = -pts[:, -1]
    return np.transpose(intrinsics @ np.transpose(pts))[:, :2]


def get_proj_corners(depth_shape, center, radius, cam_intrinsics=nocs_real_cam_intrinsics):
    """nocs_data_process.py:129-143: 2-D window (rows, cols; inclusive) of the ball's axis-aligned box."""
    radius = max(radius, 0.05)
    lo, hi = center - np.ones(3) * radius, center + np.ones(3) * radius
    aabb = np.zeros((8, 3))
    cr = np.stack([lo, hi])
    for i in range(8):                                    # nocs_utils.py:65-77
        x, y, z = (i % 4) // 2, i // 4, i % 2
        aabb[i] = (cr[x, 0], cr[y, 1], cr[z, 2])
    height, width = depth_shape
    pc = project(aabb, cam_intrinsics).astype(np.int32)[:, [1, 0]]
    pc[:, 0] = height - pc[:, 0]
    c2 = np.stack([np.min(pc, axis=0), np.max(pc, axis=0)], axis=0)
    c2[0, :] = np.maximum(c2[0, :], 0)
    c2[1, :] = np.minimum(c2[1, :], np.array([height - 1, width - 1]))
    return c2


def crop_ball_from_depth_image(depth, mask, center, radius, cam_intrinsics=nocs_real_cam_intrinsics, num_points=None,
                               device=None, perm=None, return_info=False, _depth=0):
    """nocs_data_process.py:148-164 on the device -> (pts [num_points or n, 3] float64, obj_mask [..] int32), both CUDA."""
    if not (depth.is_cuda and depth.dtype == torch.float32 and depth.is_contiguous()):
        raise _lib.CaptraError("crop_ball_from_depth_image: depth must be a contiguous CUDA float32 [H,W] tensor (no CPU path)")
    dev = depth.device
    H, W = depth.shape
    center = np.asarray(center, dtype=np.float64).reshape(3)
    c2 = get_proj_corners((H, W), center, radius, cam_intrinsics)
    r0, c0, r1, c1 = int(c2[0, 0]), int(c2[0, 1]), int(c2[1, 0]), int(c2[1, 1])
    nrows, ncols = max(r1 - r0 + 1, 0), max(c1 - c0 + 1, 0)
    npx = max(nrows * ncols, 1)
    i32 = torch.int32
    level = torch.empty(npx, dtype=torch.uint8, device=dev)
    hist = torch.empty(max(nrows, 1) * 11, dtype=i32, device=dev)
    row_off = torch.empty(max(nrows, 1), dtype=i32, device=dev)
    meta = torch.empty(3, dtype=i32, device=dev)
    pts = torch.empty(npx, 3, dtype=torch.float64, device=dev)
    pmask = torch.empty(npx, dtype=i32, device=dev)
    pix = torch.empty(npx, dtype=i32, device=dev)
    kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(cam_intrinsics, dtype=np.float64)))   # nocs_utils.py:7
    window = (ctypes.c_int * 4)(r0, c0, r1, c1)
    if mask is not None and not (mask.is_cuda and mask.dtype == i32 and mask.is_contiguous() and mask.shape == depth.shape):
        raise _lib.CaptraError("crop_ball_from_depth_image: mask must be a contiguous CUDA int32 [H,W] tensor")
    s = _lib.stream_ptr(dev)
    L = _lib.load()
    _lib.call("crop_select[%dx%d]" % (nrows, ncols), L.captra_crop_select, H, W, depth.data_ptr(),
              mask.data_ptr() if mask is not None else None, window, kinv.ctypes.data_as(ctypes.c_void_p),
              center.ctypes.data_as(ctypes.c_void_p), float(radius), 1 if num_points is not None else 0,
              level.data_ptr(), hist.data_ptr(), row_off.data_ptr(), meta.data_ptr(), pts.data_ptr(), pmask.data_ptr(), pix.data_ptr(),
              s, device=dev)
    n, thr, stop = (int(v) for v in meta.cpu())                      # the one synchronisation of the crop
    if n == 0:
        if num_points is None:
            return (pts[:0], pmask[:0]) if not return_info else (pts[:0], pmask[:0], {"n": 0, "level": stop})
        if _depth >= 30:
            raise _lib.CaptraError("crop_ball_from_depth_image: no depth measurement near the object")
        return crop_ball_from_depth_image(depth, mask, center, radius * 1.2, cam_intrinsics, num_points, device, perm,
                                          return_info, _depth + 1)            # nocs_data_process.py:159-160
    if num_points is None:
        count, sel_t, fps_idx = n, None, None
    else:
        n_t = n
        while n_t < num_points:                                      # :105-106 idx = concat(idx, idx)
            n_t *= 2
        sel_t = None
        m = n_t
        if n_t > 5 * num_points:                                     # data_utils.py:146-148
            if perm is None:
                perm = torch.randperm(n_t, device=dev)
            sel_t = torch.as_tensor(perm, device=dev)[:5 * num_points].to(torch.int64).contiguous()
            m = 5 * num_points
        cloud = torch.empty(1, m, 3, dtype=torch.float32, device=dev)
        _lib.call("crop_subset[n=%d,m=%d]" % (n, m), L.captra_crop_subset, n, m, sel_t.data_ptr() if sel_t is not None else None,
                  pts.data_ptr(), cloud.data_ptr(), s, device=dev)
        fps_idx, _ = fused_ops.fps_gather(cloud, num_points)         # pointnet_utils.py:124 (starts at index 0)
        count = num_points
    out_pts = torch.empty(count, 3, dtype=torch.float64, device=dev)
    out_mask = torch.empty(count, dtype=i32, device=dev)
    out_idx = torch.empty(count, dtype=torch.int64, device=dev)
    out_pix = torch.empty(count, dtype=i32, device=dev)
    _lib.call("crop_gather[n=%d,m=%d]" % (n, count), L.captra_crop_gather, n, count,
              fps_idx.data_ptr() if fps_idx is not None else None, sel_t.data_ptr() if sel_t is not None else None,
              pts.data_ptr(), pmask.data_ptr(), pix.data_ptr(), out_pts.data_ptr(), out_mask.data_ptr(), out_idx.data_ptr(),
              out_pix.data_ptr(), s, device=dev)
    if return_info:
        return out_pts, out_mask, {"n": n, "level": stop, "take_all": thr == 10, "idx": out_idx, "pix": out_pix, "window": (r0, c0, r1, c1)}
    return out_pts, out_mask
