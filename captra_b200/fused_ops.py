"""Thin Python fronts for the fused entry points of include/captra_ops.h section 2.
Point-major tensors in, point-major tensors out; every call is one or two kernel launches."""
import ctypes
import os

import torch

from . import _lib

_f32, _i32 = torch.float32, torch.int32


def fps_gather(xyz, npoint):
    """xyz [B,N,3] -> (idx [B,npoint] int32, new_xyz [B,npoint,3]); pointnet_utils.py:225-226."""
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=_i32, device=xyz.device)
    new_xyz = torch.empty(B, npoint, 3, dtype=_f32, device=xyz.device)
    # registers (one CTA, or a cluster above 8192 points) hold the running distances up to 32768 points; the streaming
    # kernel beyond that (or with CAPTRA_FPS_CLUSTER=0, an A/B knob) keeps them in a scratch tensor
    need_temp = N > 32768 or (N > 8192 and os.environ.get("CAPTRA_FPS_CLUSTER") == "0")
    temp = torch.full((B, N), 1e10, dtype=_f32, device=xyz.device) if need_temp else None
    _lib.call("fps_gather[B=%d,N=%d,M=%d]" % (B, N, npoint), _lib.load().captra_fps_gather, B, N, npoint,
              _lib.ptr(xyz, _f32, "xyz"), temp.data_ptr() if temp is not None else None, idx.data_ptr(), new_xyz.data_ptr(),
              _lib.stream_ptr(xyz.device), device=xyz.device)
    return idx, new_xyz


def ball_query_multi(radii, nsamples, xyz, new_xyz):
    """One scan for all radii of an MSG layer (pointnet_utils.py:228-233) -> list of idx [B,S,K_r]."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    nr = len(radii)
    # one zeroed allocation for all radii (empty balls keep their zeros, pointnet2_utils.py:261): one fill, not nr
    flat = torch.zeros(B * S * sum(int(k) for k in nsamples), dtype=_i32, device=xyz.device)
    outs, off = [], 0
    for k in nsamples:
        outs.append(flat[off:off + B * S * int(k)].view(B, S, int(k)))
        off += B * S * int(k)
    ra = (ctypes.c_float * nr)(*[float(r) for r in radii])
    ka = (ctypes.c_int * nr)(*[int(k) for k in nsamples])
    pa = (ctypes.c_void_p * nr)(*[o.data_ptr() for o in outs])
    _lib.call("ball_query_multi[B=%d,N=%d,S=%d,K=%s]" % (B, N, S, "/".join(map(str, nsamples))),
              _lib.load().captra_ball_query_multi, B, N, S, nr, ra, ka, _lib.ptr(new_xyz, _f32, "new_xyz"),
              _lib.ptr(xyz, _f32, "xyz"), pa, _lib.stream_ptr(xyz.device), device=xyz.device)
    return outs


def fps_ball_query(xyz, npoint, radii, nsamples):
    """FPS and the multi-radius ball query around its picks as one pipelined call (captra_fps_ball_query):
    xyz [B,N,3] -> (new_xyz [B,npoint,3], [idx [B,npoint,K_r] int32 per radius]); N <= 8192."""
    B, N, _ = xyz.shape
    nr = len(radii)
    tot = B * npoint * sum(int(k) for k in nsamples)
    flat = torch.zeros(tot + B, dtype=_i32, device=xyz.device)      # one fill: the index lists (empty balls keep zeros) + the progress counters
    outs, off = [], 0
    for k in nsamples:
        outs.append(flat[off:off + B * npoint * int(k)].view(B, npoint, int(k)))
        off += B * npoint * int(k)
    progress = flat[tot:]
    fps_idx = torch.empty(B, npoint, dtype=_i32, device=xyz.device)
    new_xyz = torch.empty(B, npoint, 3, dtype=_f32, device=xyz.device)
    ra = (ctypes.c_float * nr)(*[float(r) for r in radii])
    ka = (ctypes.c_int * nr)(*[int(k) for k in nsamples])
    pa = (ctypes.c_void_p * nr)(*[o.data_ptr() for o in outs])
    _lib.call("fps_ball_query[B=%d,N=%d,M=%d,K=%s]" % (B, N, npoint, "/".join(map(str, nsamples))), _lib.load().captra_fps_ball_query,
              B, N, npoint, _lib.ptr(xyz, _f32, "xyz"), fps_idx.data_ptr(), new_xyz.data_ptr(), nr, ra, ka, pa, progress.data_ptr(),
              _lib.stream_ptr(xyz.device), device=xyz.device)
    return new_xyz, outs


def ball_query_group(radius, nsample, xyz, new_xyz, features):
    """QueryAndGroup without the concat (pointnet2_utils.py:290-296): xyz [B,N,3], new_xyz [B,M,3], features [B,C,N]
    -> (idx [B,M,K] int32, grouped [B,C,M,K]); one C-ABI call (captra_ball_query_group)."""
    B, N, _ = xyz.shape
    M, C = new_xyz.shape[1], features.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=_i32, device=xyz.device)
    out = torch.empty(B, C, M, nsample, dtype=_f32, device=xyz.device)
    _lib.call("ball_query_group[B=%d,N=%d,M=%d,K=%d,C=%d]" % (B, N, M, nsample, C), _lib.load().captra_ball_query_group,
              B, N, M, C, float(radius), int(nsample), _lib.ptr(new_xyz, _f32, "new_xyz"), _lib.ptr(xyz, _f32, "xyz"),
              _lib.ptr(features, _f32, "features"), idx.data_ptr(), out.data_ptr(), _lib.stream_ptr(xyz.device), device=xyz.device)
    return idx, out


def three_nn_interpolate_pm(unknown, known, feats_pm, out=None, col_off=0, nn=None, return_nn=False):
    """unknown [B,n,3], known [B,m,3], feats_pm [B,m,C] -> out [B,n,C] (or into out[..., col_off:]);
    3-NN + inverse-distance weights + interpolation (pointnet_utils.py:284-289).  `nn` = (idx, weight)
    from an earlier call on the same coordinates skips the search."""
    B, n, _ = unknown.shape
    m, C = known.shape[1], feats_pm.shape[2]
    dev = unknown.device
    if nn is None:
        idx = torch.empty(B, n, 3, dtype=_i32, device=dev)
        w = torch.empty(B, n, 3, dtype=_f32, device=dev)
        uptr = _lib.ptr(unknown, _f32, "unknown")
    else:
        idx, w = nn
        uptr = None
    if out is None:
        out = torch.empty(B, n, C, dtype=_f32, device=dev)
    _lib.call("three_nn_interpolate[B=%d,n=%d,m=%d,C=%d%s]" % (B, n, m, C, ",reuse" if nn is not None else ""),
              _lib.load().captra_three_nn_interpolate,
              B, C, n, m, uptr, _lib.ptr(known, _f32, "known"),
        _lib.ptr(feats_pm, _f32, "feats"), _lib.ptr(out, _f32, "out"), None, idx.data_ptr(), w.data_ptr(),
        1, out.shape[-1], col_off, _lib.stream_ptr(dev), device=dev)
    return (out, (idx, w)) if return_nn else out
