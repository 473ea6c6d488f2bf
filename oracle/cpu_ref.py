"""ctypes front-end for oracle/cpu_ref.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see the header of cpu_ref.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product path
(captra_b200/) never does.

Signatures mirror the reference's Python layer (network/models/pointnet_lib/pointnet2_utils.py):
the wrappers allocate outputs the way the reference's autograd Functions do (zeros for
ball_query :261, 1e10 temp for FPS :27) so a test can call oracle and product identically.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(quiet=True):
    """(Re)build libcpu_ref.so and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcpu_ref.so")
        src = os.path.join(_HERE, "cpu_ref.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.ref_num_threads.restype = ctypes.c_int
        _LIB.ref_opt_n_threads.restype = ctypes.c_int
    return _LIB


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ci32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))


def opt_n_threads(n):
    return lib().ref_opt_n_threads(int(n))


def ball_query(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:244-264: xyz [B,N,3], new_xyz [B,M,3] -> idx [B,M,nsample] int32."""
    xyz, new_xyz = _c32(xyz), _c32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib().ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(idx))
    return idx


def grouping_operation(features, idx):
    """pointnet2_utils.py:195-215: features [B,C,N], idx [B,M,K] -> [B,C,M,K]."""
    features, idx = _c32(features), _ci32(idx)
    B, C, N = features.shape
    _, M, K = idx.shape
    out = np.empty((B, C, M, K), dtype=np.float32)
    lib().ref_group_points(B, C, N, M, K, _fp(features), _ip(idx), _fp(out))
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, idx = _c32(grad_out), _ci32(idx)
    B, C, M, K = grad_out.shape
    g = np.zeros((B, C, N), dtype=np.float32)
    lib().ref_group_points_grad(B, C, N, M, K, _fp(grad_out), _ip(idx), _fp(g))
    return g


def gather_operation(features, idx):
    """pointnet2_utils.py:40-62: features [B,C,N], idx [B,M] -> [B,C,M]."""
    features, idx = _c32(features), _ci32(idx)
    B, C, N = features.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    lib().ref_gather_points(B, C, N, M, _fp(features), _ip(idx), _fp(out))
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, idx = _c32(grad_out), _ci32(idx)
    B, C, M = grad_out.shape
    g = np.zeros((B, C, N), dtype=np.float32)
    lib().ref_gather_points_grad(B, C, N, M, _fp(grad_out), _ip(idx), _fp(g))
    return g


def furthest_point_sample(xyz, npoint, temp=None, return_temp=False):
    """pointnet2_utils.py:10-30: xyz [B,N,3] -> idx [B,npoint] int32 (temp starts at 1e10)."""
    xyz = _c32(xyz)
    B, N, _ = xyz.shape
    idx = np.empty((B, npoint), dtype=np.int32)
    temp = np.full((B, N), 1e10, dtype=np.float32) if temp is None else _c32(temp).copy()
    lib().ref_furthest_point_sampling(B, N, npoint, _fp(xyz), _fp(temp), _ip(idx))
    return (idx, temp) if return_temp else idx


def three_nn(unknown, known, sqrt=True):
    """pointnet2_utils.py:110-134: returns (sqrt(dist2), idx); sqrt=False gives raw dist2."""
    unknown, known = _c32(unknown), _c32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.empty((B, n, 3), dtype=np.float32)
    idx = np.empty((B, n, 3), dtype=np.int32)
    lib().ref_three_nn(B, n, m, _fp(unknown), _fp(known), _fp(dist2), _ip(idx))
    return (np.sqrt(dist2) if sqrt else dist2), idx


def knn(k, unknown, known, sqrt=True):
    """pointnet2_utils.py:78-104."""
    assert k <= 200
    unknown, known = _c32(unknown), _c32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.empty((B, n, k), dtype=np.float32)
    idx = np.empty((B, n, k), dtype=np.int32)
    lib().ref_knn(B, n, m, k, _fp(unknown), _fp(known), _fp(dist2), _ip(idx))
    return (np.sqrt(dist2) if sqrt else dist2), idx


def three_interpolate(features, idx, weight):
    """pointnet2_utils.py:144-170: features [B,C,m], idx/weight [B,n,3] -> [B,C,n]."""
    features, idx, weight = _c32(features), _ci32(idx), _c32(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), dtype=np.float32)
    lib().ref_three_interpolate(B, C, m, n, _fp(features), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _c32(grad_out), _ci32(idx), _c32(weight)
    B, C, n = grad_out.shape
    g = np.zeros((B, C, m), dtype=np.float32)
    lib().ref_three_interpolate_grad(B, C, n, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def interp_weights(dist):
    """pointnet_utils.py:285-287 (fp32): w = (1/(d+1e-8)) / sum."""
    dist = np.asarray(dist, dtype=np.float32)
    recip = np.float32(1.0) / (dist + np.float32(1e-8))
    norm = recip.sum(axis=2, keepdims=True, dtype=np.float32)
    return (recip / norm).astype(np.float32)


def fps_rule_reference(xyz, npoint):
    """Independent statement of the FPS tie rule: pick argmax temp with ties
    broken by (bitrev(k mod block), k).  Pure numpy; used to cross-check the thread-level
    emulation in cpu_ref.c on small tie-heavy clouds."""
    xyz = _c32(xyz)
    B, N, _ = xyz.shape
    block = opt_n_threads(N)
    bits = block.bit_length() - 1
    k = np.arange(N)
    t = k % block
    rev = np.zeros(N, dtype=np.int64)
    for i in range(bits):
        rev |= ((t >> i) & 1) << (bits - 1 - i)
    order_key = rev * (N + 1) + k  # smaller is preferred
    out = np.zeros((B, npoint), dtype=np.int32)
    for b in range(B):
        temp = np.full(N, 1e10, dtype=np.float32)
        old = 0
        for j in range(1, npoint):
            d = xyz[b] - xyz[b, old]
            dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
            t0 = (dy * dy).astype(np.float32)
            # fmaf emulation in float64 is exact for the product, single rounding on the sum
            t1 = (dx.astype(np.float64) * dx + t0).astype(np.float32)
            dd = (dz.astype(np.float64) * dz + t1).astype(np.float32)
            temp = np.fmin(dd, temp)
            mx = temp.max()
            cand = np.nonzero(temp == mx)[0]
            old = int(cand[np.argmin(order_key[cand])])
            out[b, j] = old
    return out
