"""ctypes binding of libcaptra_ops.so -- the ONLY compute path of this package.

There is deliberately no fallback: if the library is missing or an op is handed a non-CUDA
tensor the call raises.  (The CPU restatement lives under oracle/ and is test infrastructure.)
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAPTRA_LIB_PATH") or os.path.join(_PKG, "libcaptra_ops.so")   # override: A/B kernel variants

_lib = None

c_int, c_i64, c_float, c_void_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes; every entry returns int status.  Must list every symbol of
# include/captra_ops.h (tests/test_abi.py checks the header against this table).
_P = c_void_p
SIGNATURES = {
    "ball_query_kernel_launcher_fast": [c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P],
    "group_points_kernel_launcher_fast": [c_int] * 5 + [_P, _P, _P, _P],
    "group_points_grad_kernel_launcher_fast": [c_int] * 5 + [_P, _P, _P, _P],
    "gather_points_kernel_launcher_fast": [c_int] * 4 + [_P, _P, _P, _P],
    "gather_points_grad_kernel_launcher_fast": [c_int] * 4 + [_P, _P, _P, _P],
    "furthest_point_sampling_kernel_launcher": [c_int] * 3 + [_P, _P, _P, _P],
    "three_nn_kernel_launcher_fast": [c_int] * 3 + [_P, _P, _P, _P, _P],
    "knn_kernel_launcher_fast": [c_int] * 4 + [_P, _P, _P, _P, _P],
    "three_interpolate_kernel_launcher_fast": [c_int] * 4 + [_P, _P, _P, _P, _P],
    "three_interpolate_grad_kernel_launcher_fast": [c_int] * 4 + [_P, _P, _P, _P, _P],
    "captra_ball_query_multi": [c_int] * 4 + [_P, _P, _P, _P, _P, _P],
    "captra_ball_query_group": [c_int] * 4 + [c_float, c_int] + [_P] * 6,
    "captra_fps_ball_query": [c_int] * 3 + [_P, _P, _P, c_int, _P, _P, _P, _P, _P],
    "captra_fps_gather": [c_int] * 3 + [_P, _P, _P, _P, _P],
    "captra_three_nn_interpolate": [c_int] * 4 + [_P] * 7 + [c_int, c_i64, c_int, _P],
    "captra_mlp_pack": [_P, c_int, _P, _P],
    "captra_sa_mlp_max": [c_int] * 5 + [_P] * 7 + [c_i64, c_int, c_int, _P],
    "captra_sa_mlp_max_pre": [c_int] * 5 + [_P, _P, _P, c_i64, _P, _P, _P, _P, _P, c_i64, c_int, c_int, _P],
    "captra_point_mlp": [c_i64, _P, c_i64, c_int, _P, c_i64, c_int, c_int, _P, _P, _P, c_i64, c_int, c_int, c_int, _P],
    "captra_group_norm_affine": [c_int] * 4 + [_P, c_i64, _P, _P, c_float, _P, _P, _P],
    "captra_point_mlp_affine": [c_i64, _P, c_i64, c_int, _P, _P, c_int, _P, _P, _P, c_i64, c_int, c_int, _P],
    "captra_point_mlp_gnstats": [c_i64, _P, c_i64, c_int, _P, _P, c_int, _P, _P, _P, c_i64, c_int, _P, c_int, _P],
    "captra_group_norm_relu_rows": [c_i64, c_int, c_int, _P, c_i64, _P, _P, _P],
    "captra_group_norm_finalize": [c_int] * 4 + [_P, _P, _P, c_float, _P, _P, _P],
    "captra_debug_tc_timestamps": [_P, c_int],
    "captra_f16_overflow_flag": [c_int],
    "captra_debug_umma_gemm": [c_int, c_int, _P, _P, _P, c_int, _P],
    "captra_procrustes_rot3": [c_i64, _P, _P, _P],
    "captra_procrustes_rot2": [c_i64, _P, _P, _P],
    "captra_part_fit_track": [c_int] * 3 + [_P] * 5 + [c_int] + [_P] * 6,
    "captra_canonicalize": [c_int] * 3 + [_P] * 9,
    "captra_coord_head_post": [c_int] * 4 + [_P, c_i64, _P, c_i64, _P, _P, _P, _P],
    "captra_rot_head_post": [c_int] * 4 + [_P, c_i64, _P, _P, _P, _P, _P],
    "captra_track_eval": [c_int] * 5 + [_P] * 14 + [c_int, _P],
    "captra_crop_select": [c_int, c_int, _P, _P, _P, _P, _P, ctypes.c_double, c_int] + [_P] * 8,
    "captra_crop_subset": [c_int, c_int, _P, _P, _P, _P],
    "captra_crop_gather": [c_int, c_int] + [_P] * 10,
    "captra_part_fit_st": [c_int] * 3 + [_P, _P] + [_P] + [c_i64] * 4 + [_P] + [c_i64] * 4 + [_P, _P, c_int, _P, _P, _P, _P, _P],
}
OTHER_SYMBOLS = ["captra_last_error", "captra_abi_version", "captra_launch_count", "captra_mlp_pack_bytes"]


class CaptraError(RuntimeError):
    pass


def load():
    """Load libcaptra_ops.so (raises if it has not been built: run `python -m captra_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CaptraError(
            "libcaptra_ops.so not found at %s -- build it with `python -m captra_b200.build` "
            "(there is no CPU/eager fallback in this package)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.captra_last_error.restype = ctypes.c_char_p
    lib.captra_abi_version.restype = c_int
    lib.captra_launch_count.restype = c_i64
    lib.captra_mlp_pack_bytes.restype = c_i64
    lib.captra_mlp_pack_bytes.argtypes = [_P, c_int]
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        raise CaptraError("%s failed (status %d): %s" % (what, status, load().captra_last_error().decode()))


# --- optional per-call device timing (bench.py's roofline pass) ---------------------------------
# When PROFILE is a list, call() brackets every C-ABI call with CUDA events on the current stream
# and appends (tag, start_event, end_event).  Off (None) in normal operation: zero overhead.
PROFILE = None


def call(tag, fn, *args, device=None):
    """Invoke a C-ABI entry point, raise on a non-zero status, optionally time it.  The launch goes to `device`
    (the tensors' device), not to whatever device happens to be current."""
    if device is not None and device.index is not None and device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            return call(tag, fn, *args, device=None)
    if PROFILE is None:
        status = fn(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(device))
        status = fn(*args)
        e1.record(torch.cuda.current_stream(device))
        PROFILE.append((tag, e0, e1))
    if status != 0:
        raise CaptraError("%s failed (status %d): %s" % (tag, status, load().captra_last_error().decode()))


def launch_count():
    return int(load().captra_launch_count())


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (raises otherwise -- no silent copies)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CaptraError("%s must be a CUDA tensor (captra_b200 has no CPU path)" % name)
    if not t.is_contiguous():
        raise CaptraError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise CaptraError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.data_ptr()
