"""GPU parity tests for the pointnet_lib ops, through the
reference-shaped Python API -> C ABI -> sm_100a kernels.

Bars: indices bit-exact; copies bit-exact; three_interpolate bit-exact (same FMA order);
atomic-add adjoints to 1e-5 (summation order is unspecified in the reference too).
Three-way check: product == oracle (CPU restatement) == reference CUDA kernels (oracle/_ref).
"""
import numpy as np
import pytest
import torch

from captra_b200 import synthetic

pytestmark = pytest.mark.gpu


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


@pytest.fixture(scope="module")
def futils(cuda):
    from captra_b200.pointnet_lib import pointnet2_utils
    return pointnet2_utils


@pytest.fixture(scope="module")
def refcu(cuda):
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    return ref_cuda


CLOUDS = {
    "surface4096": lambda: synthetic.batch_surface_box(3, 4096, seed=0)[0],
    "uniform5000": lambda: synthetic.batch_uniform(2, 5000, seed=1),       # n not a power of two
    "tiled4096": lambda: synthetic.batch_tiled(2, 4096, 1500, seed=2),     # exact duplicates
    "small100": lambda: synthetic.batch_uniform(4, 100, seed=3),
    "sa2_512": lambda: synthetic.batch_surface_box(4, 512, seed=4)[0],
    "n1": lambda: synthetic.batch_uniform(2, 1, seed=5),
    "n33": lambda: synthetic.batch_tiled(2, 33, 5, seed=6),
}


@pytest.mark.parametrize("name,m", [("surface4096", 512), ("uniform5000", 300), ("tiled4096", 2000),
                                    ("small100", 100), ("sa2_512", 128), ("n1", 3), ("n33", 33)])
def test_fps_bit_exact(name, m, futils, oracle, refcu, cuda):
    pts = CLOUDS[name]()
    want = oracle.furthest_point_sample(pts, m)
    got = futils.furthest_point_sample(dev(pts, cuda), m)
    assert got.dtype == torch.int32 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    ref = refcu.furthest_point_sample(dev(pts, cuda), m)
    assert np.array_equal(ref.cpu().numpy(), want), "oracle disagrees with the reference kernel"


def test_fps_large_n_stream_path_and_temp(oracle, refcu, cuda):
    # crop FPS: up to 20480 points -> 4096 (data_utils.py:146-158); also checks the scratch
    # `temp` the reference leaves behind
    pts = synthetic.batch_surface_box(1, 20480, seed=7)[0]
    from captra_b200 import pointnet2_cuda
    x = dev(pts, cuda)
    for n, m in ((20480, 1024), (9000, 500), (8192, 700)):
        xs = x[:, :n].contiguous()
        temp = torch.full((1, n), 1e10, device=cuda)
        idx = torch.empty(1, m, dtype=torch.int32, device=cuda)
        pointnet2_cuda.furthest_point_sampling_wrapper(1, n, m, xs, temp, idx)
        want, wtemp = oracle.furthest_point_sample(pts[:, :n], m, return_temp=True)
        assert np.array_equal(idx.cpu().numpy(), want)
        assert np.array_equal(temp.cpu().numpy(), wtemp)
    ridx, rtemp = refcu.furthest_point_sample(x[:, :9000].contiguous(), 500, return_temp=True)
    want, wtemp = oracle.furthest_point_sample(pts[:, :9000], 500, return_temp=True)
    assert np.array_equal(ridx.cpu().numpy(), want) and np.array_equal(rtemp.cpu().numpy(), wtemp)


@pytest.mark.parametrize("name,m,radius,k", [
    ("surface4096", 512, 0.05, 32), ("surface4096", 512, 0.1, 64), ("surface4096", 512, 0.2, 128),
    ("sa2_512", 128, 0.2, 64), ("sa2_512", 128, 0.4, 128), ("uniform5000", 77, 0.15, 16),
    ("tiled4096", 100, 0.1, 48), ("small100", 100, 0.3, 200), ("n1", 1, 1.0, 4)])
def test_ball_query_bit_exact(name, m, radius, k, futils, oracle, refcu, cuda):
    pts = CLOUDS[name]()
    ctr = np.ascontiguousarray(pts[:, oracle.furthest_point_sample(pts, m)[0]])
    ctr[0, 0] += 50.0  # one empty ball: row must stay zero
    want = oracle.ball_query(radius, k, pts, ctr)
    got = futils.ball_query(radius, k, dev(pts, cuda), dev(ctr, cuda))
    assert got.dtype == torch.int32
    assert np.array_equal(got.cpu().numpy(), want)
    assert (got[0, 0] == 0).all()
    ref = refcu.ball_query(radius, k, dev(pts, cuda), dev(ctr, cuda))
    assert np.array_equal(ref.cpu().numpy(), want), "oracle disagrees with the reference kernel"


def test_ball_query_multi_radius_equals_single(oracle, cuda):
    import ctypes
    from captra_b200 import _lib
    L = _lib.load()
    pts = CLOUDS["surface4096"]()
    ctr = np.ascontiguousarray(pts[:, :512])
    x, c = dev(pts, cuda), dev(ctr, cuda)
    radii, ks = [0.05, 0.1, 0.2], [32, 64, 128]
    outs = [torch.zeros(3, 512, k, dtype=torch.int32, device=cuda) for k in ks]
    ra = (ctypes.c_float * 3)(*radii)
    ka = (ctypes.c_int * 3)(*ks)
    pa = (ctypes.c_void_p * 3)(*[o.data_ptr() for o in outs])
    _lib.check(L.captra_ball_query_multi(3, 4096, 512, 3, ra, ka, c.data_ptr(), x.data_ptr(), pa,
                                         _lib.stream_ptr()), "ball_query_multi")
    for r, k, o in zip(radii, ks, outs):
        assert np.array_equal(o.cpu().numpy(), oracle.ball_query(r, k, pts, ctr))


def test_group_gather_bit_exact(futils, oracle, refcu, cuda):
    rng = np.random.default_rng(0)
    for (B, C, N, M, K) in ((2, 6, 4096, 512, 32), (2, 323, 512, 128, 64), (1, 5, 301, 7, 3)):
        feats = rng.normal(size=(B, C, N)).astype(np.float32)
        idx = rng.integers(0, N, size=(B, M, K)).astype(np.int32)
        want = oracle.grouping_operation(feats, idx)
        got = futils.grouping_operation(dev(feats, cuda), dev(idx, cuda))
        assert np.array_equal(got.cpu().numpy(), want)
        assert np.array_equal(refcu.grouping_operation(dev(feats, cuda), dev(idx, cuda)).cpu().numpy(), want)
        sidx = np.ascontiguousarray(idx[:, :, 0])
        want = oracle.gather_operation(feats, sidx)
        got = futils.gather_operation(dev(feats, cuda), dev(sidx, cuda))
        assert np.array_equal(got.cpu().numpy(), want)
        assert np.array_equal(refcu.gather_operation(dev(feats, cuda), dev(sidx, cuda)).cpu().numpy(), want)


def test_three_nn_interpolate_bit_exact(futils, oracle, refcu, cuda):
    for (B, n, m, C, seed) in ((2, 4096, 512, 128, 0), (2, 512, 128, 256, 1), (1, 77, 2, 5, 2), (1, 300, 3000, 4, 3)):
        unk = synthetic.batch_uniform(B, n, seed=seed)
        kn = synthetic.batch_uniform(B, m, seed=seed + 10)
        kn[:, 1:3] = kn[:, 0:1] if m > 3 else kn[:, 1:3]  # duplicated known points: tie order
        wd, wi = oracle.three_nn(unk, kn, sqrt=False)
        gd, gi = futils.three_nn(dev(unk, cuda), dev(kn, cuda))
        assert np.array_equal(gi.cpu().numpy(), wi)
        assert np.array_equal(gd.cpu().numpy(), np.sqrt(wd))
        rd, ri = refcu.three_nn(dev(unk, cuda), dev(kn, cuda))
        assert np.array_equal(ri.cpu().numpy(), wi) and np.array_equal(rd.cpu().numpy(), wd)
        feats = np.random.default_rng(seed).normal(size=(B, C, m)).astype(np.float32)
        w = oracle.interp_weights(np.sqrt(wd)) if m >= 3 else np.full((B, n, 3), 1 / 3, np.float32)
        want = oracle.three_interpolate(feats, wi, w)
        got = futils.three_interpolate(dev(feats, cuda), dev(wi, cuda), dev(w, cuda))
        assert np.array_equal(got.cpu().numpy(), want)
        ref = refcu.three_interpolate(dev(feats, cuda), dev(wi, cuda), dev(w, cuda))
        assert np.array_equal(ref.cpu().numpy(), want), "oracle disagrees with the reference kernel"


def test_knn_bit_exact(futils, oracle, refcu, cuda):
    unk = synthetic.batch_uniform(2, 333, seed=0)
    kn = synthetic.batch_uniform(2, 500, seed=1)
    for k in (1, 3, 16, 200):
        wd, wi = oracle.knn(k, unk, kn, sqrt=False)
        gd, gi = futils.knn(k, dev(unk, cuda), dev(kn, cuda))
        assert np.array_equal(gi.cpu().numpy(), wi)
        assert np.array_equal(gd.cpu().numpy(), np.sqrt(wd))
        rd, ri = refcu.knn(k, dev(unk, cuda), dev(kn, cuda))
        assert np.array_equal(ri.cpu().numpy(), wi) and np.array_equal(rd.cpu().numpy(), wd)


def test_backward_ops(futils, oracle, refcu, cuda):
    rng = np.random.default_rng(0)
    B, C, N, M, K, n = 2, 16, 512, 64, 8, 700
    x = dev(rng.normal(size=(B, C, N)).astype(np.float32), cuda).requires_grad_(True)
    idx = rng.integers(0, N, size=(B, M, K)).astype(np.int32)
    g = rng.normal(size=(B, C, M, K)).astype(np.float32)
    futils.grouping_operation(x, dev(idx, cuda)).backward(dev(g, cuda))
    np.testing.assert_allclose(x.grad.cpu().numpy(), oracle.grouping_operation_grad(g, idx, N), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(refcu.grouping_operation_grad(dev(g, cuda), dev(idx, cuda), N).cpu().numpy(),
                               oracle.grouping_operation_grad(g, idx, N), rtol=1e-5, atol=1e-5)
    x.grad = None
    sidx = np.ascontiguousarray(idx[:, :, 0])
    g2 = rng.normal(size=(B, C, M)).astype(np.float32)
    futils.gather_operation(x, dev(sidx, cuda)).backward(dev(g2, cuda))
    np.testing.assert_allclose(x.grad.cpu().numpy(), oracle.gather_operation_grad(g2, sidx, N), rtol=1e-5, atol=1e-5)
    x.grad = None
    idx3 = rng.integers(0, N, size=(B, n, 3)).astype(np.int32)
    w = rng.random((B, n, 3)).astype(np.float32)
    g3 = rng.normal(size=(B, C, n)).astype(np.float32)
    futils.three_interpolate(x, dev(idx3, cuda), dev(w, cuda)).backward(dev(g3, cuda))
    np.testing.assert_allclose(x.grad.cpu().numpy(), oracle.three_interpolate_grad(g3, idx3, w, N), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(refcu.three_interpolate_grad(dev(g3, cuda), dev(idx3, cuda), dev(w, cuda), N).cpu().numpy(),
                               oracle.three_interpolate_grad(g3, idx3, w, N), rtol=1e-5, atol=1e-5)


def test_rejects_cpu_tensors_and_missing_lib(futils, cuda):
    from captra_b200._lib import CaptraError
    with pytest.raises(CaptraError):
        futils.furthest_point_sample(torch.zeros(1, 10, 3), 2)


def test_empty_batch_and_zero_sizes(futils, cuda):
    assert futils.ball_query(0.1, 4, torch.zeros(0, 10, 3, device=cuda), torch.zeros(0, 2, 3, device=cuda)).shape == (0, 2, 4)
    assert futils.furthest_point_sample(torch.zeros(0, 10, 3, device=cuda), 4).shape == (0, 4)
    out = futils.grouping_operation(torch.zeros(2, 0, 10, device=cuda), torch.zeros(2, 3, 4, dtype=torch.int32, device=cuda))
    assert out.shape == (2, 0, 3, 4)


@pytest.mark.parametrize("b,c,n,m,k", [
    (4, 70, 1000, 256, 64),      # rows staged in shared memory, 16-byte stores, ragged last channel block
    (3, 33, 1001, 250, 63),      # same with L % 4 != 0 and unaligned rows: scalar variant
    (1, 3, 16384, 4096, 64),     # BASELINE cfg5 level 1 shape (one cloud): slot range split over CTAs
    (2, 130, 30000, 2048, 128),  # rows of 120 KB: one channel per CTA
])
def test_group_points_large_calls_bit_exact(b, c, n, m, k, futils, oracle, cuda):
    """Big group_points calls take the shared-memory-staged gather (csrc/group_gather.cu)."""
    rng = np.random.default_rng(b * 1000 + c)
    feats = rng.normal(size=(b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, k)).astype(np.int32)
    idx[0, 0, :4] = [0, n - 1, 0, n - 1]
    want = oracle.grouping_operation(feats, idx)
    got = futils.grouping_operation(dev(feats, cuda), dev(idx, cuda))
    assert np.array_equal(got.cpu().numpy(), want)
    idx1 = np.ascontiguousarray(idx.reshape(b, m * k))
    got1 = futils.gather_operation(dev(feats, cuda), dev(idx1, cuda))
    assert np.array_equal(got1.cpu().numpy(), want.reshape(b, c, m * k))


@pytest.mark.parametrize("kind,n,m,radius,k", [
    ("surface", 16384, 512, 0.05, 64),      # BASELINE cfg5 level 1: sparse balls, bitmap path
    ("uniform", 12000, 300, 0.08, 32),
    ("surface", 16384, 256, 0.3, 64),       # dense balls: > 512 candidates -> early-exit scan fallback
    ("tiled", 10000, 200, 0.06, 48),        # exact duplicates
    ("surface", 20480, 128, 0.02, 16),      # many empty balls
    ("surface", 2048, 300, 0.1, 64),        # smallest cloud on the grid path (2 bitmap words per lane)
    ("uniform", 4096, 1024, 0.1, 64),       # BASELINE cfg5 level 2
    ("uniform", 70001, 97, 0.03, 40),       # bitmap too big for four warps per CTA: one-warp variant
    ("surface", 5000, 77, 0.12, 200),       # nsample > hits for most balls, odd sizes
    ("uniform", 2048, 2048, 0.2, 64),       # every point a centroid: crowded cells split over several work items
    ("uniform", 16384, 4096, 0.05, 64),     # BASELINE cfg5 level 1, all 4096 centroids
    ("clump", 6000, 1500, 0.05, 32),        # > 512 candidates around most cells: per-centroid search inside the cell kernel
    ("same", 4096, 64, 0.05, 16),           # one cell holds everything
])
def test_ball_query_grid_path_bit_exact(kind, n, m, radius, k, futils, oracle, refcu, cuda):
    """Large clouds take the binned search (csrc/ball_query.cu); same hits, same order as the scan."""
    if kind == "surface":
        pts = np.stack([synthetic.surface_box(n, np.random.default_rng(7 + i))[0] for i in range(2)])
    elif kind == "uniform":
        pts = synthetic.batch_uniform(2, n, seed=3)
    elif kind == "clump":
        pts = (0.12 * synthetic.batch_uniform(2, n, seed=5)).astype(np.float32)
    elif kind == "same":
        pts = np.full((2, n, 3), 0.25, np.float32)
    else:
        pts = synthetic.batch_tiled(2, n, 1234, seed=4)
    ctr = np.ascontiguousarray(pts[:, ::n // m][:, :m]).copy()
    ctr[0, 0] += 50.0                      # far outside the bounding box: empty
    ctr[1, 1] = pts[1].min(0) - 0.9 * radius   # just outside the box corner
    ctr[0, 2, 0] = np.nan                  # NaN centroid: no hits
    want = oracle.ball_query(radius, k, pts, ctr)
    got = futils.ball_query(radius, k, dev(pts, cuda), dev(ctr, cuda))
    assert np.array_equal(got.cpu().numpy(), want)
    ref = refcu.ball_query(radius, k, dev(pts, cuda), dev(ctr, cuda))
    assert np.array_equal(ref.cpu().numpy(), want)


@pytest.mark.parametrize("N,M,K,r,C", [(16384, 1024, 64, 0.05, 3), (4096, 512, 32, 0.1, 3), (1000, 100, 16, 0.2, 3),
                                        (4096, 256, 64, 0.2, 64), (8192, 300, 64, 0.02, 6)])
def test_ball_query_group_fused(N, M, K, r, C, cuda, oracle):
    """captra_ball_query_group (QueryAndGroup in one call; the query warp writes the grouped rows itself for few
    channels) == ball_query then group_points, bit for bit, incl. empty balls (r=0.02: isolated outliers) and a NaN centroid."""
    from captra_b200 import fused_ops, synthetic
    from captra_b200.pointnet_lib import pointnet2_utils as futils
    pts = synthetic.batch_surface_box(2, N, seed=N + K)[0]
    x = torch.from_numpy(pts).to(cuda)
    ctr = x[:, torch.randperm(N, generator=torch.Generator().manual_seed(1))[:M]].contiguous()
    ctr[0, 0] = 10.0                              # far away: empty ball
    ctr[1, 1, 0] = float("nan")
    feats = x.transpose(1, 2).contiguous() if C == 3 else torch.randn(2, C, N, device=cuda)
    idx, grouped = fused_ops.ball_query_group(r, K, x, ctr, feats)
    want_idx = oracle.ball_query(r, K, pts, ctr.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert torch.equal(idx, futils.ball_query(r, K, x, ctr))
    assert torch.equal(grouped, futils.grouping_operation(feats, idx))
    assert (idx[0, 0] == 0).all() and (idx[1, 1] == 0).all()


@pytest.mark.parametrize("n,m,kind", [(16384, 1024, "surface"), (12000, 700, "tiled"), (32768, 300, "uniform"), (20480, 4096, "tiled"),
                                       (40000, 200, "uniform")])
def test_fps_cluster_path(n, m, kind, oracle, cuda):
    """Clouds above 8192 points: a thread-block cluster per cloud, records exchanged through distributed shared memory
    (n <= 32768), the streaming kernel beyond; no caller scratch (fused_ops.fps_gather).  Bit-exact incl. exhausted
    duplicates (tiled: as the data loader pads small crops, nocs_data_process.py:105-106)."""
    from captra_b200 import fused_ops
    if kind == "surface":
        pts = synthetic.batch_surface_box(3, n, seed=n)[0]
    elif kind == "tiled":
        pts = synthetic.batch_tiled(2, n, 3000, seed=n)
    else:
        pts = synthetic.batch_uniform(2, n, seed=n)
    idx, new_xyz = fused_ops.fps_gather(dev(pts, cuda), m)
    want = oracle.furthest_point_sample(pts, m)
    assert np.array_equal(idx.cpu().numpy(), want)
    assert np.array_equal(new_xyz.cpu().numpy(), np.take_along_axis(pts, want.astype(np.int64)[..., None], 1))


def test_backward_ops_are_deterministic_and_correctly_rounded(cuda):
    """The three training-only adjoints accumulate in 64-bit fixed point (csrc/det_accum.cuh): bit-identical from run to
    run whatever the atomics' interleaving (the reference's fp32 atomicAdd is not: group_points_gpu.cu:8-25,
    sampling_gpu.cu:46-63, interpolate_gpu.cu:192-214), and equal to the exactly summed gradient rounded to fp32."""
    from captra_b200 import pointnet2_cuda as P
    gen = torch.Generator().manual_seed(0)
    B, C, N, M, K, n = 2, 32, 64, 256, 64, 5000                       # few destinations, thousands of colliding contributions each
    idx = torch.randint(0, N, (B, M, K), generator=gen, dtype=torch.int32).to(cuda)
    g = (torch.randn(B, C, M, K, generator=gen) * torch.logspace(-6, 2, C).view(1, C, 1, 1)).to(cuda)   # 8 decades of magnitudes
    runs = []
    for _ in range(3):
        out = torch.zeros(B, C, N, device=cuda)
        P.group_points_grad_wrapper(B, C, N, M, K, g, idx, out)
        runs.append(out)
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    exact = torch.zeros(B, C, N, dtype=torch.float64, device=cuda)
    exact.scatter_add_(2, idx.view(B, 1, M * K).expand(-1, C, -1).long(), g.double().view(B, C, M * K))
    # the fixed-point grid is 2^-41 of the call's max|g| per contribution
    tol = float(g.abs().max()) * M * K * 2.0 ** -41
    assert float((runs[0].double() - exact).abs().max()) <= tol + float(exact.abs().max()) * 2.0 ** -24
    idx3 = torch.randint(0, N, (B, n, 3), generator=gen, dtype=torch.int32).to(cuda)
    w = torch.rand(B, n, 3, generator=gen).to(cuda)
    g3 = torch.randn(B, C, n, generator=gen).to(cuda)
    runs = []
    for _ in range(3):
        out = torch.zeros(B, C, N, device=cuda)
        P.three_interpolate_grad_wrapper(B, C, n, N, g3, idx3, w, out)
        runs.append(out)
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    exact = torch.zeros(B, C, N, dtype=torch.float64, device=cuda)
    for j in range(3):
        exact.scatter_add_(2, idx3[..., j].view(B, 1, n).expand(-1, C, -1).long(), (g3 * w[..., j].view(B, 1, n)).double())
    assert float((runs[0].double() - exact).abs().max()) <= 1e-5 * float(exact.abs().max())
    sidx = torch.randint(0, N, (B, M), generator=gen, dtype=torch.int32).to(cuda)
    g2 = torch.randn(B, C, M, generator=gen).to(cuda)
    a, b = torch.zeros(B, C, N, device=cuda), torch.zeros(B, C, N, device=cuda)
    P.gather_points_grad_wrapper(B, C, N, M, g2, sidx, a)
    P.gather_points_grad_wrapper(B, C, N, M, g2, sidx, b)
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,N,M,radii,ks,kind", [(32, 4096, 512, [0.05, 0.1, 0.2], [32, 64, 128], "surface"), (5, 512, 128, [0.2, 0.4], [64, 128], "surface"),
                                                   (3, 1000, 77, [0.15], [16], "uniform"), (2, 4096, 2000, [0.05, 0.3], [8, 32], "tiled"),
                                                   (70, 300, 33, [0.1, 0.2, 0.3, 0.4], [4, 8, 16, 32], "uniform")])
def test_fps_ball_query_pipeline(B, N, M, radii, ks, kind, oracle, cuda):
    """captra_fps_ball_query: FPS and the multi-radius ball query overlapped through a progress counter and a
    programmatic dependent launch -- bit-identical to the two separate ops (and the oracle), also when replayed
    from a CUDA graph and with more ball-query blocks than fit the GPU at once (B = 70)."""
    from captra_b200 import fused_ops
    if kind == "surface":
        pts = synthetic.batch_surface_box(B, N, seed=B + N)[0]
    elif kind == "tiled":
        pts = synthetic.batch_tiled(B, N, 1500, seed=N)
    else:
        pts = synthetic.batch_uniform(B, N, seed=N)
    x = dev(pts, cuda)
    new_xyz, idxs = fused_ops.fps_ball_query(x, M, radii, ks)
    _, want_xyz = fused_ops.fps_gather(x, M)
    want_idx = fused_ops.ball_query_multi(radii, ks, x, want_xyz)
    assert torch.equal(new_xyz, want_xyz)
    for got, want, r, k in zip(idxs, want_idx, radii, ks):
        assert torch.equal(got, want)
        if B <= 5:
            assert np.array_equal(got.cpu().numpy(), oracle.ball_query(r, k, pts, want_xyz.cpu().numpy()))
    # replayed from a graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fused_ops.fps_ball_query(x, M, radii, ks)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gx, gi = fused_ops.fps_ball_query(x, M, radii, ks)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(gx, want_xyz) and all(torch.equal(a, b) for a, b in zip(gi, want_idx))
