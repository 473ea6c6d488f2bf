"""GPU unit tests for the hand-written tcgen05 primitives (csrc/tc_common.cuh) through the
captra_debug_umma_gemm doorway: descriptor encodings, operand layout, TMEM load mapping, and the
3xTF32 split arithmetic vs an fp64 GEMM."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, W, terms, cuda):
    from captra_b200 import _lib
    K, N = A.shape[1], W.shape[0]
    D = torch.full((128, N), float("nan"), device=cuda)
    _lib.call("debug_umma_gemm", _lib.load().captra_debug_umma_gemm, K, N, A.data_ptr(), W.data_ptr(), D.data_ptr(),
              terms, _lib.stream_ptr(cuda))
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("K,N", [(8, 16), (16, 64), (64, 128), (32, 256), (64, 208), (96, 96)])
def test_umma_gemm_layout_and_3xtf32(K, N, cuda):
    gen = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=gen).to(cuda)
    W = torch.randn(N, K, generator=gen).to(cuda)
    want = (A.double() @ W.double().t())
    scale = (A.abs().double() @ W.abs().double().t())
    d3 = _run(A, W, 3, cuda).double()
    err3 = ((d3 - want).abs() / scale).max().item()
    assert err3 < 2e-6, "3xTF32 relative error %.3g (layout or split is wrong)" % err3
    d1 = _run(A, W, 1, cuda).double()
    err1 = ((d1 - want).abs() / scale).max().item()
    assert 1e-5 < err1 < 2e-3, "single-pass TF32 error %.3g outside the expected band" % err1


def test_umma_gemm_identity_mapping(cuda):
    # W = I picks out columns of A: catches row/column permutations in the TMEM load path
    A = torch.arange(128 * 32, dtype=torch.float32, device=cuda).reshape(128, 32) / 64.0
    W = torch.eye(32, device=cuda)
    D = _run(A, W, 3, cuda)
    assert torch.equal(D, A)
