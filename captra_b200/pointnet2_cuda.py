"""Drop-in for the reference's pybind extension module `pointnet2_cuda`.

Same ten functions, same positional signatures and return values as
network/models/pointnet_lib/src/pointnet2_api.cpp:10-25 (wrappers in ball_query.cpp:14,
group_points.cpp:11,25, sampling.cpp:11,24,38, interpolate.cpp:14,26,39,55); each forwards the
raw device pointers and torch's current CUDA stream to the C ABI in include/captra_ops.h.
`captra_b200.install_dropin()` registers this module as top-level `pointnet2_cuda`, which is what
the reference's pointnet2_utils.py:7 imports.

Differences, all deliberate: every tensor is checked for CUDA / contiguity / dtype (the
reference checks only ball_query, ball_query.cpp:10-17) and a CUDA launch error raises
CaptraError instead of exit(-1) (e.g. sampling_gpu.cu:248-252).
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_f32, _i32 = torch.float32, torch.int32


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    L = _lib.load()
    check(L.ball_query_kernel_launcher_fast(
        b, n, m, float(radius), nsample, ptr(new_xyz_tensor, _f32, "new_xyz"), ptr(xyz_tensor, _f32, "xyz"),
        ptr(idx_tensor, _i32, "idx"), stream_ptr(xyz_tensor.device)), "ball_query_wrapper")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    L = _lib.load()
    check(L.group_points_kernel_launcher_fast(
        b, c, n, npoints, nsample, ptr(points_tensor, _f32, "points"), ptr(idx_tensor, _i32, "idx"),
        ptr(out_tensor, _f32, "out"), stream_ptr(points_tensor.device)), "group_points_wrapper")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    L = _lib.load()
    check(L.group_points_grad_kernel_launcher_fast(
        b, c, n, npoints, nsample, ptr(grad_out_tensor, _f32, "grad_out"), ptr(idx_tensor, _i32, "idx"),
        ptr(grad_points_tensor, _f32, "grad_points"), stream_ptr(grad_out_tensor.device)), "group_points_grad_wrapper")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    L = _lib.load()
    check(L.gather_points_kernel_launcher_fast(
        b, c, n, npoints, ptr(points_tensor, _f32, "points"), ptr(idx_tensor, _i32, "idx"),
        ptr(out_tensor, _f32, "out"), stream_ptr(points_tensor.device)), "gather_points_wrapper")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    L = _lib.load()
    check(L.gather_points_grad_kernel_launcher_fast(
        b, c, n, npoints, ptr(grad_out_tensor, _f32, "grad_out"), ptr(idx_tensor, _i32, "idx"),
        ptr(grad_points_tensor, _f32, "grad_points"), stream_ptr(grad_out_tensor.device)), "gather_points_grad_wrapper")
    return 1


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    L = _lib.load()
    check(L.furthest_point_sampling_kernel_launcher(
        b, n, m, ptr(points_tensor, _f32, "points"), ptr(temp_tensor, _f32, "temp"),
        ptr(idx_tensor, _i32, "idx"), stream_ptr(points_tensor.device)), "furthest_point_sampling_wrapper")
    return 1


def knn_wrapper(b, n, m, k, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    L = _lib.load()
    check(L.knn_kernel_launcher_fast(
        b, n, m, k, ptr(unknown_tensor, _f32, "unknown"), ptr(known_tensor, _f32, "known"),
        ptr(dist2_tensor, _f32, "dist2"), ptr(idx_tensor, _i32, "idx"), stream_ptr(unknown_tensor.device)), "knn_wrapper")


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    L = _lib.load()
    check(L.three_nn_kernel_launcher_fast(
        b, n, m, ptr(unknown_tensor, _f32, "unknown"), ptr(known_tensor, _f32, "known"),
        ptr(dist2_tensor, _f32, "dist2"), ptr(idx_tensor, _i32, "idx"), stream_ptr(unknown_tensor.device)), "three_nn_wrapper")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    L = _lib.load()
    check(L.three_interpolate_kernel_launcher_fast(
        b, c, m, n, ptr(points_tensor, _f32, "points"), ptr(idx_tensor, _i32, "idx"),
        ptr(weight_tensor, _f32, "weight"), ptr(out_tensor, _f32, "out"), stream_ptr(points_tensor.device)),
        "three_interpolate_wrapper")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    L = _lib.load()
    check(L.three_interpolate_grad_kernel_launcher_fast(
        b, c, n, m, ptr(grad_out_tensor, _f32, "grad_out"), ptr(idx_tensor, _i32, "idx"),
        ptr(weight_tensor, _f32, "weight"), ptr(grad_points_tensor, _f32, "grad_points"),
        stream_ptr(grad_out_tensor.device)), "three_interpolate_grad_wrapper")
