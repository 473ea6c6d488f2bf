"""Mirror of pose_utils/pose_fit.py: part_fit_st_no_ransac / filter_model_valid, same signatures.

The tracker calls part_fit_st_no_ransac once per frame (networks.py:227-228).  The reference
builds a one-hot mask, runs ~10 small torch kernels and, for symmetric categories, a CPU SVD;
here the whole fit is ONE kernel launch (csrc/pose_fit.cu: captra_part_fit_st) that reads the
labels and the transposed [B,P,3,N] views in place through their strides.
"""
import torch

from .. import _lib
from .procrustes import transform_pts_mask


def filter_model_valid(model, valid):
    """pose_fit.py:26-35: drop parts whose scale/translation/rotation contain NaN or Inf."""
    for key in ("scale", "translation", "rotation"):
        tmp = model[key] if key == "scale" else model[key].sum((-1, -2))
        valid = torch.logical_and(valid, torch.isfinite(tmp))
    return valid


def _fusable(labels, source, target, rotation, given_scale):
    ts = [labels, source, target] + [t for t in (rotation, given_scale) if t is not None]
    if not all(t.is_cuda for t in ts):
        raise _lib.CaptraError("part_fit_st_no_ransac: tensors must be on a CUDA device (no CPU path)")
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts if t.is_floating_point()):
        return False
    return (source.dim() == 4 and source.shape == target.shape and source.shape[-1] == 3
            and source.dtype == torch.float32 and target.dtype == torch.float32
            and labels.dim() == 2 and labels.dtype == torch.int64)


def part_fit_st_no_ransac(labels, source, target, rotation, cfg, given_scale=None):
    """pose_fit.py:38-53.  labels [B,N] int64, source/target [B,P,N,3] (any strides),
    rotation [B,P,3,3] -> ({'rotation','scale' [B,P],'translation' [B,P,3,1]}, valid [B,P] bool).
    As in the reference, the returned rotation is the *input* rotation (the sym-refined one is
    only used internally for s and t)."""
    num_parts = cfg["num_parts"]
    if not _fusable(labels, source, target, rotation, given_scale):
        eye = torch.cat([torch.eye(num_parts), torch.zeros(2, num_parts)], dim=0).to(labels.device)
        mask = eye[labels, ].transpose(-1, -2)
        valid = mask.sum(dim=-1) > 3
        _, scale, translation = transform_pts_mask(source, target, mask.unsqueeze(-1), mask.unsqueeze(-1),
                                                   given_scale=given_scale, rotation=rotation, sym=cfg["sym"])
        model = {"rotation": rotation, "scale": scale, "translation": translation}
        return model, filter_model_valid(model, valid)

    B, P, N, _ = source.shape
    assert P == num_parts, "source has %d parts, cfg says %d" % (P, num_parts)
    labels = labels.contiguous()
    rot = rotation.to(torch.float32).contiguous() if rotation is not None else None
    gs = given_scale.to(torch.float32).contiguous() if given_scale is not None else None
    dev = source.device
    scale = torch.empty(B, P, dtype=torch.float32, device=dev)
    translation = torch.empty(B, P, 3, 1, dtype=torch.float32, device=dev)
    valid = torch.empty(B, P, dtype=torch.uint8, device=dev)
    rot_used = torch.empty(B, P, 3, 3, dtype=torch.float32, device=dev) if rotation is None else None
    ss, ts = source.stride(), target.stride()
    _lib.call("part_fit_st[B=%d,P=%d,N=%d]" % (B, P, N), _lib.load().captra_part_fit_st,
        B, P, N, labels.data_ptr(), None,
        source.data_ptr(), ss[0], ss[1], ss[2], ss[3],
        target.data_ptr(), ts[0], ts[1], ts[2], ts[3],
        rot.data_ptr() if rot is not None else None, gs.data_ptr() if gs is not None else None,
        1 if cfg["sym"] else 0, scale.data_ptr(), translation.data_ptr(), valid.data_ptr(),
        rot_used.data_ptr() if rot_used is not None else None, _lib.stream_ptr(dev), device=dev)
    model = {"rotation": rotation if rotation is not None else rot_used, "scale": scale, "translation": translation}
    return model, valid.bool()
