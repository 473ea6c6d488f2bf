"""ctypes front-end for oracle/_ref/libpointnet2_ref.so -- the REFERENCE's own CUDA kernels
(network/models/pointnet_lib/src/*_gpu.cu, compiled unmodified by oracle/Makefile) behind the
extern "C" shim oracle/ref_shim.cu.  TEST INFRASTRUCTURE ONLY: the GPU tests use it to pin
oracle/cpu_ref.c against the real reference and bench.py may time it as the "reference GPU
path"; the product never loads it.  Takes/returns torch CUDA tensors; outputs are allocated
the way the reference's pointnet2_utils.py does."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libpointnet2_ref.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(PATH)
    return _LIB


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32, device=xyz.device)
    lib().refcu_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def grouping_operation(features, idx):
    B, C, N = features.shape
    _, M, K = idx.shape
    out = torch.empty(B, C, M, K, device=features.device)
    lib().refcu_group_points(B, C, N, M, K, _p(features), _p(idx), _p(out), _s())
    return out


def gather_operation(features, idx):
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, device=features.device)
    lib().refcu_gather_points(B, C, N, M, _p(features), _p(idx), _p(out), _s())
    return out


def furthest_point_sample(xyz, npoint, return_temp=False):
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, device=xyz.device)
    lib().refcu_furthest_point_sampling(B, N, npoint, _p(xyz), _p(temp), _p(idx), _s())
    return (idx, temp) if return_temp else idx


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, 3, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    lib().refcu_three_nn(B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def knn(k, unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, k, device=unknown.device)
    idx = torch.empty(B, n, k, dtype=torch.int32, device=unknown.device)
    lib().refcu_knn(B, n, m, k, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_interpolate(features, idx, weight):
    B, C, m = features.shape
    n = idx.shape[1]
    out = torch.empty(B, C, n, device=features.device)
    lib().refcu_three_interpolate(B, C, m, n, _p(features), _p(idx), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    g = torch.zeros(B, C, m, device=grad_out.device)
    lib().refcu_three_interpolate_grad(B, C, n, m, _p(grad_out), _p(idx), _p(weight), _p(g), _s())
    return g


def grouping_operation_grad(grad_out, idx, N):
    B, C, M, K = grad_out.shape
    g = torch.zeros(B, C, N, device=grad_out.device)
    lib().refcu_group_points_grad(B, C, N, M, K, _p(grad_out), _p(idx), _p(g), _s())
    return g


def gather_operation_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    g = torch.zeros(B, C, N, device=grad_out.device)
    lib().refcu_gather_points_grad(B, C, N, M, _p(grad_out), _p(idx), _p(g), _s())
    return g
