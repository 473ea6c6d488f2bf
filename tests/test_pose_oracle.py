"""CPU: pin oracle/pose_ref.py (fp64 numpy restatement) against tests/golden/procrustes.npz,
i.e. against outputs of the reference's own pose_utils code (torch fp32 + LAPACK)."""
import os

import numpy as np
import pytest

from oracle import pose_ref as PR

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "procrustes.npz"))
TOL = dict(rtol=1e-4, atol=2e-6)  # the north star's fp32 pose tolerance (1e-4 rel)


@pytest.mark.parametrize("name", ["rigid_sym", "arti", "arti3"])
def test_part_fit_matches_reference(name):
    g = lambda k: GOLD[name + "/" + k]
    labels, src, rot = g("labels"), g("source"), g("rotation")
    P = src.shape[1]
    tgt = np.repeat(g("cam")[:, None], P, 1)
    model, valid = PR.part_fit_st_no_ransac(labels, src, tgt, rot, {"num_parts": P, "sym": bool(g("sym"))})
    assert np.array_equal(valid, g("fit_valid"))
    np.testing.assert_allclose(model["scale"], g("fit_scale"), **TOL)
    np.testing.assert_allclose(model["translation"], g("fit_translation"), **TOL)
    eye = np.concatenate([np.eye(P), np.zeros((2, P))], 0)
    mask = np.swapaxes(eye[labels], -1, -2)[..., None]
    R, s, t = PR.transform_pts_mask(src, tgt, mask, mask, rotation=None, sym=False)
    np.testing.assert_allclose(R, g("full_R"), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s, g("full_s"), **TOL)
    np.testing.assert_allclose(t, g("full_t"), rtol=1e-4, atol=1e-5)
    R2, t2 = PR.transform_pts_2d_mask(src[..., [0, 2]], (tgt @ rot)[..., [0, 2]], mask)
    np.testing.assert_allclose(R2, g("rot2d"), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(t2, g("trans2d"), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(PR.scale_pts_mask(src * mask, tgt * mask, mask), g("scale_mask"), **TOL)
    np.testing.assert_allclose(PR.translate_pts_mask(np.swapaxes(src, -1, -2), np.swapaxes(tgt, -1, -2), mask),
                               g("translate_mask"), rtol=1e-4, atol=1e-5)


def test_unmasked_and_raw_rotations_match_reference():
    R, s, t = PR.transform_pts_batch(GOLD["batch/source"], GOLD["batch/target"])
    np.testing.assert_allclose(R, GOLD["batch/R"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s, GOLD["batch/s"], **TOL)
    np.testing.assert_allclose(t, GOLD["batch/t"], rtol=1e-4, atol=1e-5)
    R3 = PR.rotate_pts_batch(GOLD["rot3/src"].astype(np.float64), GOLD["rot3/tgt"].astype(np.float64))
    np.testing.assert_allclose(R3, GOLD["rot3/R"], rtol=1e-4, atol=2e-5)
    assert np.allclose(np.linalg.det(R3), 1.0, atol=1e-6)
    R2 = PR.rotate_pts_2d_batch(GOLD["rot2/src"].astype(np.float64), GOLD["rot2/tgt"].astype(np.float64))
    np.testing.assert_allclose(R2, GOLD["rot2/R"], rtol=1e-4, atol=2e-5)
