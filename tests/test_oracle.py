"""CPU tests: pin oracle/cpu_ref against (i) fixtures produced by the reference's own Python code
(tests/golden/make_golden.py), (ii) independent numpy statements of each op's rule.  The pin
against the reference's CUDA kernels themselves (oracle/_ref) runs on the GPU box
(the `refcu` comparisons of tests/test_ops_gpu.py: test_fps_bit_exact, test_ball_query_bit_exact, test_group_gather_bit_exact, test_three_nn_interpolate_bit_exact, test_knn_bit_exact, test_backward_ops)."""
import os

import numpy as np
import pytest

from captra_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_opt_n_threads_matches_integer_log2(oracle):
    # cuda_utils.h:10-14 uses log(n)/log(2); fps.cu uses an integer log2.  Exhaustive to 2^22.
    assert oracle.lib().ref_check_opt_n_threads(1 << 22) == 0
    assert [oracle.opt_n_threads(n) for n in (1, 2, 3, 128, 512, 1000, 4096, 5000, 20480)] == \
        [1, 2, 2, 128, 512, 512, 1024, 1024, 1024]


def test_fps_emulation_matches_tie_rule(oracle):
    # thread-level emulation (cpu_ref.c) vs the closed-form rule on clouds
    # made of exact duplicates, where almost every round is decided by the tie-break
    for n, unique, m in ((1000, 300, 400), (256, 40, 100), (700, 699, 64), (64, 7, 20)):
        pts = synthetic.batch_tiled(2, n, unique, seed=n)
        a = oracle.furthest_point_sample(pts, m)
        b = oracle.fps_rule_reference(pts, m)
        assert np.array_equal(a, b), (n, unique, m)


def test_fps_known_example_bitrev(oracle):
    # block=32; equal maxima at k=7,14,18,29 -> winner 18 (bitrev 9), not 7
    pts = np.zeros((1, 32, 3), np.float32)
    pts[0, [7, 14, 18, 29], 0] = 1.0
    idx = oracle.furthest_point_sample(pts, 2)
    assert idx[0].tolist() == [0, 18]


def test_fps_basic_properties(oracle):
    pts = synthetic.batch_uniform(3, 777, seed=1)
    idx, temp = oracle.furthest_point_sample(pts, 100, return_temp=True)
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 100  # distinct points while distances are > 0
    # temp holds min sq-distance to the first m-1 picks
    sel = pts[0, idx[0, :-1]]
    d = ((pts[0][:, None, :] - sel[None]) ** 2).sum(-1).min(1)
    np.testing.assert_allclose(temp[0], d, rtol=1e-5, atol=1e-7)


def _naive_ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = np.zeros((B, M, nsample), np.int32)
    r2 = np.float32(radius) * np.float32(radius)
    for b in range(B):
        d = new_xyz[b][:, None, :] - xyz[b][None]
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        t = (dy * dy).astype(np.float32)
        t = (dx.astype(np.float64) * dx + t).astype(np.float32)
        d2 = (dz.astype(np.float64) * dz + t).astype(np.float32)
        for m in range(M):
            hits = np.nonzero(d2[m] < r2)[0][:nsample]
            if len(hits):
                out[b, m, :] = hits[0]
                out[b, m, :len(hits)] = hits
    return out


def test_ball_query_rule(oracle):
    pts, _ = synthetic.batch_surface_box(2, 600, seed=2)
    ctr = pts[:, :50].copy()
    ctr[0, 3] = 10.0  # empty ball -> row stays zero
    for r, k in ((0.05, 8), (0.2, 16), (0.4, 64)):
        got = oracle.ball_query(r, k, pts, ctr)
        assert np.array_equal(got, _naive_ball_query(r, k, pts, ctr))
    assert (oracle.ball_query(0.1, 8, pts, ctr)[0, 3] == 0).all()


def test_three_nn_rule(oracle):
    unk = synthetic.batch_uniform(2, 200, seed=3)
    kn = synthetic.batch_uniform(2, 64, seed=4)
    d2, idx = oracle.three_nn(unk, kn, sqrt=False)
    full = ((unk[:, :, None, :] - kn[:, None]) ** 2).sum(-1)
    order = np.argsort(full, axis=-1, kind="stable")[:, :, :3]
    assert np.array_equal(idx, order.astype(np.int32))
    np.testing.assert_allclose(d2, np.take_along_axis(full, order, -1), rtol=1e-5, atol=1e-7)
    # ties: duplicated known points -> earlier index first
    kn2 = np.concatenate([kn[:, :5], kn[:, :5]], 1)
    _, idx2 = oracle.three_nn(unk, kn2)
    assert (idx2[..., 0] < 5).all() and np.array_equal(idx2[..., 1], idx2[..., 0] + 5)
    # m < 3 -> inf distances, index 0 (interpolate_gpu.cu:102: sentinels never replaced)
    d1, i1 = oracle.three_nn(unk, kn[:, :2], sqrt=False)
    assert np.isinf(d1[..., 2]).all() and (i1[..., 2] == 0).all()


def test_knn_matches_three_nn_and_sort(oracle):
    unk = synthetic.batch_uniform(1, 50, seed=5)
    kn = synthetic.batch_uniform(1, 80, seed=6)
    d3, i3 = oracle.three_nn(unk, kn, sqrt=False)
    dk, ik = oracle.knn(7, unk, kn, sqrt=False)
    assert np.array_equal(ik[..., :3], i3) and np.array_equal(dk[..., :3], d3)
    assert (np.diff(dk, axis=-1) >= 0).all()


def test_index_ops_match_reference_python(oracle):
    g = np.load(os.path.join(GOLD, "ops_index.npz"))
    assert np.array_equal(oracle.grouping_operation(g["feats"], g["gidx"]), g["grouped"])
    assert np.array_equal(oracle.gather_operation(g["feats"], g["sidx"]), g["gathered"])
    # reference CPU three_interpolate sums in a different order -> float tolerance
    np.testing.assert_allclose(oracle.three_interpolate(g["feats"], g["idx3"], g["w3"]), g["interp"],
                               rtol=1e-5, atol=1e-6)


def test_grad_ops_are_adjoints(oracle):
    rng = np.random.default_rng(0)
    B, C, N, M, K = 2, 3, 50, 7, 4
    x = rng.normal(size=(B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, M, K)).astype(np.int32)
    g = rng.normal(size=(B, C, M, K)).astype(np.float32)
    lhs = (oracle.grouping_operation(x, idx) * g).sum()
    rhs = (x * oracle.grouping_operation_grad(g, idx, N)).sum()
    np.testing.assert_allclose(lhs, rhs, rtol=1e-4)
    sidx = idx[:, :, 0].copy()
    g2 = rng.normal(size=(B, C, M)).astype(np.float32)
    np.testing.assert_allclose((oracle.gather_operation(x, sidx) * g2).sum(),
                               (x * oracle.gather_operation_grad(g2, sidx, N)).sum(), rtol=1e-4)
    idx3 = rng.integers(0, N, size=(B, 9, 3)).astype(np.int32)
    w = rng.random((B, 9, 3)).astype(np.float32)
    g3 = rng.normal(size=(B, C, 9)).astype(np.float32)
    np.testing.assert_allclose((oracle.three_interpolate(x, idx3, w) * g3).sum(),
                               (x * oracle.three_interpolate_grad(g3, idx3, w, N)).sum(), rtol=1e-4)
