"""A `pointnet2_cuda` module (the reference's pybind API, pointnet2_api.cpp:10-25) backed by the REFERENCE's own CUDA
kernels in oracle/_ref/libpointnet2_ref.so.  TEST / BASELINE INFRASTRUCTURE ONLY: with this registered as
sys.modules['pointnet2_cuda'] and oracle/_ref/pyref on sys.path, the reference's whole GPU path (its kernels, its
pointnet2_utils.py, its torch modules, its CPU torch.svd) runs on the box -- bench.py times that as the
"reference GPU path" frames/s next to this library's."""
import ctypes

import torch

from . import ref_cuda


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    ref_cuda.lib().refcu_ball_query(b, n, m, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    ref_cuda.lib().refcu_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), _s())
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    ref_cuda.lib().refcu_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(grad_points), _s())
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    ref_cuda.lib().refcu_gather_points(b, c, n, npoints, _p(points), _p(idx), _p(out), _s())
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    ref_cuda.lib().refcu_gather_points_grad(b, c, n, npoints, _p(grad_out), _p(idx), _p(grad_points), _s())
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    ref_cuda.lib().refcu_furthest_point_sampling(b, n, m, _p(points), _p(temp), _p(idx), _s())
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    ref_cuda.lib().refcu_knn(b, n, m, k, _p(unknown), _p(known), _p(dist2), _p(idx), _s())


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    ref_cuda.lib().refcu_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), _s())


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    ref_cuda.lib().refcu_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    ref_cuda.lib().refcu_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(grad_points), _s())
