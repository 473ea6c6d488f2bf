"""Python fronts of the per-frame glue kernels (include/captra_ops.h section 4b, csrc/frame_glue.cu) and of
captra_part_fit_track: canonicalisation, CoordNet head post-processing, RotationRegressor head
post-processing + composition, and the tracker's pose fit -- one launch each, CUDA-graph capturable."""
import ctypes

import torch

from . import _lib

_f32, _i64 = torch.float32, torch.int64


def canonicalize(points, points_mean, rotation, translation, scale, parts=1, want_cm=False, want_dup=False):
    """networks.py:38-41 / :170-187.  points [B,3,N], points_mean [B,3,1], rotation [B*P,3,3] (or [B,P,3,3]),
    translation [B*P,3,1], scale [B*P] -> (xyz_pm [B*P,N,3], cam [B*P,3,N] or None, dup [B*P,N,6] or None)."""
    B, _, N = points.shape
    dev = points.device
    xyz_pm = torch.empty(B * parts, N, 3, dtype=_f32, device=dev)
    cm = torch.empty(B * parts, 3, N, dtype=_f32, device=dev) if want_cm else None
    dup = torch.empty(B * parts, N, 6, dtype=_f32, device=dev) if want_dup else None
    # contiguous copies (the root part's slice of a multi-part pose) must stay referenced until the launch is enqueued
    points, points_mean, rotation, translation, scale = (t.contiguous() for t in (points, points_mean, rotation, translation, scale))
    _lib.call("canonicalize[B=%d,P=%d,N=%d]" % (B, parts, N), _lib.load().captra_canonicalize, B, parts, N,
              _lib.ptr(points, _f32, "points"), _lib.ptr(points_mean, _f32, "points_mean"),
              _lib.ptr(rotation, _f32, "rotation"), _lib.ptr(translation, _f32, "translation"),
              _lib.ptr(scale, _f32, "scale"), xyz_pm.data_ptr(), cm.data_ptr() if want_cm else None,
              dup.data_ptr() if want_dup else None, _lib.stream_ptr(dev), device=dev)
    return xyz_pm, cm, dup


def coord_head_post(seg_raw, nocs_raw, B, N):
    """seg_raw [B*N, nseg], nocs_raw [B*N, 3P] (point-major head outputs) -> (labels [B,N] int64,
    nocs [B,3P,N], seg [B,nseg,N]); networks.py:44-46, model.py:458."""
    dev = seg_raw.device
    nseg, nnocs = seg_raw.shape[1], nocs_raw.shape[1]
    labels = torch.empty(B, N, dtype=_i64, device=dev)
    nocs = torch.empty(B, nnocs, N, dtype=_f32, device=dev)
    seg = torch.empty(B, nseg, N, dtype=_f32, device=dev)
    _lib.call("coord_head_post[B=%d,N=%d,S=%d,C=%d]" % (B, N, nseg, nnocs), _lib.load().captra_coord_head_post, B, N, nseg, nnocs,
              _lib.ptr(seg_raw, _f32, "seg_raw"), seg_raw.stride(0), _lib.ptr(nocs_raw, _f32, "nocs_raw"), nocs_raw.stride(0),
              labels.data_ptr(), nocs.data_ptr(), seg.data_ptr(), _lib.stream_ptr(dev), device=dev)
    return labels, nocs, seg


def rot_head_post(raws, labels, rot_prev, sym, want_rtvec=False):
    """raws: list of P tensors [B,N,D] (head p on copy p, point-major), labels [B,N] int64, rot_prev [B,P,3,3]
    -> rotation [B,P,3,3] (and the averaged prediction rtvec [B,P,3 or 9]); blocks.py:181-193,
    networks.py:127-141, part_dof_utils.py:124-141."""
    P = len(raws)
    B, N, D = raws[0].shape
    dev = raws[0].device
    rotation = torch.empty(B, P, 3, 3, dtype=_f32, device=dev)
    rtvec = torch.empty(B, P, 3 if sym else 9, dtype=_f32, device=dev) if want_rtvec else None
    ptrs = (ctypes.c_void_p * P)(*[_lib.ptr(r, _f32, "raw") for r in raws])
    rot_prev = rot_prev.contiguous()
    _lib.call("rot_head_post[B=%d,P=%d,N=%d,D=%d]" % (B, P, N, D), _lib.load().captra_rot_head_post, B, P, N, 1 if sym else 0,
              ptrs, D, _lib.ptr(labels, _i64, "labels"), _lib.ptr(rot_prev, _f32, "rot_prev"),
              rotation.data_ptr(), rtvec.data_ptr() if want_rtvec else None, _lib.stream_ptr(dev), device=dev)
    return (rotation, rtvec) if want_rtvec else rotation


def part_fit_track(labels, nocs, points, points_mean, rotation, sym, prev_scale, prev_translation):
    """networks.py:218-232 in one launch.  labels [B,N] int64, nocs [B,P,3,N], points [B,3,N], points_mean [B,3,1],
    rotation / prev pose [B,P,...] -> (scale [B,P], translation [B,P,3,1], valid [B,P] bool)."""
    B, P, _, N = nocs.shape
    dev = nocs.device
    scale = torch.empty(B, P, dtype=_f32, device=dev)
    translation = torch.empty(B, P, 3, 1, dtype=_f32, device=dev)
    valid = torch.empty(B, P, dtype=torch.uint8, device=dev)
    points, points_mean, rotation, prev_scale, prev_translation = (
        t.contiguous() for t in (points, points_mean, rotation, prev_scale, prev_translation))      # kept referenced until enqueued
    _lib.call("part_fit_st[B=%d,P=%d,N=%d]" % (B, P, N), _lib.load().captra_part_fit_track, B, P, N,
              _lib.ptr(labels, _i64, "labels"), _lib.ptr(nocs, _f32, "nocs"), _lib.ptr(points, _f32, "points"),
              _lib.ptr(points_mean, _f32, "points_mean"), _lib.ptr(rotation, _f32, "rotation"),
              1 if sym else 0, _lib.ptr(prev_scale, _f32, "prev_scale"),
              _lib.ptr(prev_translation, _f32, "prev_translation"), scale.data_ptr(), translation.data_ptr(),
              valid.data_ptr(), _lib.stream_ptr(dev), device=dev)
    return scale, translation, valid.bool()


def track_eval(gt, pose, sym, pred=None, gt_labels=None, gt_nocs=None, per_instance=False, out=None, accumulate=False):
    """Per-frame eval / loss SUMS of the tracking loop in two launches (captra_track_eval): gt / pose are part-pose
    dicts ([B,P,3,3], [B,P,3,1], [B,P]); pred = the CoordNet predictions {'seg','nocs','labels'} and gt_labels [B,N]
    int64 / gt_nocs [B,3,N] enable the mIoU / NOCS losses (model.py:546-561).  Returns sums [5P + 5] (layout:
    include/captra_ops.h), ready for an all-reduce(SUM); eval_means() turns them into the dict the reference logs.
    out + accumulate=True adds this frame to the sums `out` already holds."""
    B, P = pose["scale"].shape
    dev = pose["scale"].device
    sums = out if out is not None else torch.empty(5 * P + 5, dtype=_f32, device=dev)
    per = torch.empty(B, P, 5, dtype=_f32, device=dev) if per_instance else None
    seg = nocs = labels = None
    n = nseg = 0
    if pred is not None and (gt_labels is not None or gt_nocs is not None):
        seg, nocs, labels = pred["seg"], pred["nocs"], pred["labels"]
        nseg, n = seg.shape[1], seg.shape[2]
    scratch = torch.empty(max(B, 1) * 3, dtype=_f32, device=dev)
    keep = [t.contiguous() for t in (gt["rotation"], gt["translation"], gt["scale"], pose["rotation"], pose["translation"], pose["scale"])]
    g = lambda i: _lib.ptr(keep[i], _f32, "pose")       # `keep` holds any contiguous copy until the launch is enqueued
    _lib.call("track_eval[B=%d,P=%d,N=%d]" % (B, P, n), _lib.load().captra_track_eval, B, P, n, nseg, 1 if sym else 0,
              g(0), g(1), g(2), g(3), g(4), g(5),
              _lib.ptr(seg, _f32, "seg"), _lib.ptr(nocs, _f32, "nocs"), _lib.ptr(labels, _i64, "labels"),
              _lib.ptr(gt_labels, _i64, "gt_labels"), _lib.ptr(gt_nocs, _f32, "gt_nocs"), scratch.data_ptr(),
              per.data_ptr() if per is not None else None, sums.data_ptr(), 1 if accumulate else 0, _lib.stream_ptr(dev), device=dev)
    return (sums, per) if per_instance else sums


def eval_means(sums, num_parts):
    """sums [5P + 5] (after the all-reduce) -> the per-part means eval_part_full returns (part_dof_utils.py:54-67:
    keys sdiff_i, tdiff_i, rdiff_i, 5deg5cm_i, 10deg10cm_i) plus seg_loss / nocs_loss (loss.py:42-70,122-134)."""
    s = sums.detach().double().cpu()
    count = max(float(s[5 * num_parts]), 1.0)
    out = {}
    for p in range(num_parts):
        for q, name in enumerate(("sdiff", "tdiff", "rdiff", "5deg5cm", "10deg10cm")):
            out["%s_%d" % (name, p)] = float(s[5 * p + q]) / count
    o = 5 * num_parts
    if float(s[o + 2]) > 0:
        out["seg_loss"] = 1.0 - float(s[o + 1]) / float(s[o + 2])
    if float(s[o + 4]) > 0:
        out["nocs_loss"] = float(s[o + 3]) / max(float(s[o + 4]), 1.0)
    out["count"] = float(s[o])
    return out
