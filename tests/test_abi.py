"""CPU: the C-ABI library loads and exports every symbol include/captra_ops.h declares; the
ctypes table in captra_b200/_lib.py covers exactly those symbols; the drop-in module exposes the
reference's ten pybind functions; nothing under captra_b200/ imports the oracle."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "captra_ops.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(?:int|int64_t|const char \*)\s*\*?\s*(\w+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from captra_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libcaptra_ops.so does not export %s" % name
    table = set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    assert table == declared, (table - declared, declared - table)
    assert lib.captra_abi_version() == 1
    assert lib.captra_launch_count() == 0  # no compute has happened


def test_arg_validation_without_gpu():
    # pure host-side argument checks return an error code before any CUDA call
    from captra_b200 import _lib
    lib = _lib.load()
    rc = lib.knn_kernel_launcher_fast(1, 4, 4, 500, None, None, None, None, None)
    assert rc == 1 and b"outside 1..200" in lib.captra_last_error()
    rc = lib.captra_fps_gather(1, 1 << 23, 4, None, None, None, None, None)
    assert rc == 1 and b"2^22" in lib.captra_last_error()
    assert lib.ball_query_kernel_launcher_fast(0, 10, 10, 0.1, 4, None, None, None, None) == 0  # empty batch: no-op


def test_dropin_module_surface():
    import captra_b200
    mod = captra_b200.install_dropin()
    import pointnet2_cuda
    assert pointnet2_cuda is mod
    want = {  # pointnet2_api.cpp:10-25 with arities from the wrappers' signatures
        "ball_query_wrapper": 8, "group_points_wrapper": 8, "group_points_grad_wrapper": 8,
        "gather_points_wrapper": 7, "gather_points_grad_wrapper": 7, "furthest_point_sampling_wrapper": 6,
        "knn_wrapper": 8, "three_nn_wrapper": 7, "three_interpolate_wrapper": 8, "three_interpolate_grad_wrapper": 8}
    for name, arity in want.items():
        assert len(inspect.signature(getattr(mod, name)).parameters) == arity, name
    from pointnet_lib import pointnet2_utils as futils
    for name in ("furthest_point_sample", "gather_operation", "knn", "three_nn", "three_interpolate",
                 "grouping_operation", "ball_query", "QueryAndGroup", "GroupAll", "KNNAndGroup"):
        assert hasattr(futils, name)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_reference_python_layer_imports_against_dropin():
    """The reference's own pointnet2_utils.py (`import pointnet2_cuda as pointnet2`, :7) and
    pose_fit.py import unchanged when the drop-in is installed."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import captra_b200; captra_b200.install_dropin()\n"
        "sys.path.insert(0, '/root/reference/network/models/pointnet_lib')\n"
        "import importlib.util as u\n"
        "s = u.spec_from_file_location('ref_p2u', '/root/reference/network/models/pointnet_lib/pointnet2_utils.py')\n"
        "m = u.module_from_spec(s); s.loader.exec_module(m)\n"
        "assert m.pointnet2.__name__ == 'captra_b200.pointnet2_cuda'\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "captra_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dp, f)
                assert "cpu_ref" not in text and "libpointnet2_ref" not in text, os.path.join(dp, f)


def test_missing_library_fails_loudly(monkeypatch):
    from captra_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcaptra_ops.so")
    with pytest.raises(_lib.CaptraError):
        _lib.load()


def test_mlp_pack_bytes_is_host_only_and_shape_dispatch():
    """captra_mlp_pack_bytes is pure host arithmetic (no GPU needed): the fused tcgen05 layouts accept the chains
    the tracker keeps on chip and refuse (-1) the ones it runs layer by layer; the exact-fp32 layout takes all."""
    import ctypes
    from captra_b200 import _lib
    from captra_b200.mlp import MlpDesc
    L = _lib.load()

    def desc(cin, couts):
        d = MlpDesc()
        d.nlayers, d.cin, d.relu_last = len(couts), cin, 1
        for i, c in enumerate(couts):
            d.cout[i] = c
        return d

    fused = [(6, [64, 96, 128]), (3, [32, 32, 64]), (323, [128, 196, 256]), (128, [196, 256]), (134, [128, 128, 128]), (512, [512])]
    for cin, couts in fused:
        for impl in (0, 1, 2):
            n = L.captra_mlp_pack_bytes(ctypes.byref(desc(cin, couts)), impl)
            assert n > 0 and n % 4 == 0, (cin, couts, impl, n)
    # the fp16 pack holds two 2-byte planes per weight, the tf32 pack two 4-byte planes
    d = desc(323, [128, 196, 256])
    assert L.captra_mlp_pack_bytes(ctypes.byref(d), 2) < L.captra_mlp_pack_bytes(ctypes.byref(d), 1)
    # sa3 (515 -> 256 -> 512 -> 1024): a layer wider than 256 columns inside a chain is not fused on the tensor cores
    d = desc(515, [256, 512, 1024])
    assert L.captra_mlp_pack_bytes(ctypes.byref(d), 1) == -1 and L.captra_mlp_pack_bytes(ctypes.byref(d), 2) == -1
    assert L.captra_mlp_pack_bytes(ctypes.byref(d), 0) > 0
    assert L.captra_mlp_pack_bytes(ctypes.byref(d), 7) == -1
