"""GPU parity for the pose fit: device kernels, through the
reference-shaped Python API, vs the fp64 oracle and vs the reference's own outputs
(tests/golden/procrustes.npz).  Bar: fp32 pose within 1e-4 relative (north star)."""
import os

import numpy as np
import pytest
import torch

from captra_b200 import synthetic
from oracle import pose_ref as PR

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "procrustes.npz"))
TOL = dict(rtol=1e-4, atol=2e-6)
RTOL = dict(rtol=1e-4, atol=2e-5)


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


@pytest.fixture(scope="module")
def P(cuda):
    from captra_b200.pose_utils import procrustes
    return procrustes


@pytest.fixture(scope="module")
def PF(cuda):
    from captra_b200.pose_utils import pose_fit
    return pose_fit


@pytest.mark.parametrize("name", ["rigid_sym", "arti", "arti3"])
def test_part_fit_fused_vs_golden_and_oracle(name, P, PF, cuda):
    g = lambda k: GOLD[name + "/" + k]
    labels, src, rot, sym = g("labels"), g("source"), g("rotation"), bool(g("sym"))
    Pn = src.shape[1]
    cfg = {"num_parts": Pn, "sym": sym}
    # feed the exact views the tracker passes: [B,P,3,N].transpose(-1,-2) (networks.py:227)
    src_t = dev(np.swapaxes(src, -1, -2), cuda).transpose(-1, -2)
    tgt_t = dev(g("cam"), cuda).transpose(1, 2).unsqueeze(1).repeat(1, Pn, 1, 1).transpose(-1, -2)
    assert not src_t.is_contiguous()
    model, valid = PF.part_fit_st_no_ransac(dev(labels, cuda), src_t, tgt_t, dev(rot, cuda), cfg)
    assert valid.dtype == torch.bool and model["translation"].shape == (src.shape[0], Pn, 3, 1)
    assert np.array_equal(valid.cpu().numpy(), g("fit_valid"))
    np.testing.assert_allclose(model["scale"].cpu().numpy(), g("fit_scale"), **TOL)
    np.testing.assert_allclose(model["translation"].cpu().numpy(), g("fit_translation"), **RTOL)
    assert model["rotation"].data_ptr() == dev(rot, cuda).data_ptr() or torch.equal(model["rotation"].cpu(), torch.from_numpy(rot))
    # rotation=None -> fused 3x3 Procrustes
    tgt = np.repeat(g("cam")[:, None], Pn, 1)
    model2, valid2 = PF.part_fit_st_no_ransac(dev(labels, cuda), dev(src, cuda), dev(tgt, cuda), None,
                                              {"num_parts": Pn, "sym": False})
    np.testing.assert_allclose(model2["rotation"].cpu().numpy(), g("full_R"), **RTOL)
    np.testing.assert_allclose(model2["scale"].cpu().numpy(), g("full_s"), **TOL)
    np.testing.assert_allclose(model2["translation"].cpu().numpy(), g("full_t"), **RTOL)
    # given_scale path vs oracle
    gs = np.full(src.shape[:2], 0.3, np.float32)
    m3, _ = PF.part_fit_st_no_ransac(dev(labels, cuda), dev(src, cuda), dev(tgt, cuda), dev(rot, cuda), cfg,
                                     given_scale=dev(gs, cuda))
    o3, _ = PR.part_fit_st_no_ransac(labels, src, tgt, rot, cfg, given_scale=gs.astype(np.float64))
    np.testing.assert_allclose(m3["translation"].cpu().numpy(), o3["translation"], **RTOL)


def test_unfused_api_vs_golden(P, cuda):
    for name in ("rigid_sym", "arti"):
        g = lambda k: GOLD[name + "/" + k]
        labels, src, rot = g("labels"), g("source"), g("rotation")
        Pn = src.shape[1]
        tgt = np.repeat(g("cam")[:, None], Pn, 1)
        eye = np.concatenate([np.eye(Pn), np.zeros((2, Pn))], 0).astype(np.float32)
        mask = dev(np.swapaxes(eye[labels], -1, -2)[..., None], cuda)
        R, s, t = P.transform_pts_mask(dev(src, cuda), dev(tgt, cuda), mask, mask, rotation=None, sym=False)
        np.testing.assert_allclose(R.cpu().numpy(), g("full_R"), **RTOL)
        np.testing.assert_allclose(s.cpu().numpy(), g("full_s"), **TOL)
        np.testing.assert_allclose(t.cpu().numpy(), g("full_t"), **RTOL)
        R2, t2 = P.transform_pts_2d_mask(dev(src[..., [0, 2]], cuda), dev((tgt @ rot)[..., [0, 2]], cuda), mask)
        np.testing.assert_allclose(R2.cpu().numpy(), g("rot2d"), **RTOL)
        np.testing.assert_allclose(t2.cpu().numpy(), g("trans2d"), **RTOL)
        # sym + rotation given through the unfused composition == fused kernel
        Rs, ss, ts = P.transform_pts_mask(dev(src, cuda), dev(tgt, cuda), mask, mask, rotation=dev(rot, cuda), sym=True)
        oR, os_, ot = PR.transform_pts_mask(src, tgt, mask.cpu().numpy(), mask.cpu().numpy(), rotation=rot, sym=True)
        np.testing.assert_allclose(Rs.cpu().numpy(), oR, **RTOL)
        np.testing.assert_allclose(ss.cpu().numpy(), os_, **TOL)
        np.testing.assert_allclose(ts.cpu().numpy(), ot, **RTOL)
    R, s, t = P.transform_pts_batch(dev(GOLD["batch/source"], cuda), dev(GOLD["batch/target"], cuda))
    np.testing.assert_allclose(R.cpu().numpy(), GOLD["batch/R"], **RTOL)
    np.testing.assert_allclose(s.cpu().numpy(), GOLD["batch/s"], **TOL)
    np.testing.assert_allclose(t.cpu().numpy(), GOLD["batch/t"], **RTOL)


def test_rotation_kernels(P, cuda):
    R3 = P.rotate_pts_batch(dev(GOLD["rot3/src"], cuda), dev(GOLD["rot3/tgt"], cuda)).cpu().numpy()
    np.testing.assert_allclose(R3, GOLD["rot3/R"], **RTOL)
    assert np.allclose(np.linalg.det(R3.astype(np.float64)), 1.0, atol=1e-5)
    R2 = P.rotate_pts_2d_batch(dev(GOLD["rot2/src"], cuda), dev(GOLD["rot2/tgt"], cuda)).cpu().numpy()
    np.testing.assert_allclose(R2, GOLD["rot2/R"], **RTOL)
    # large random batch vs the fp64 oracle, incl. reflections (det M < 0) and rank-deficient M
    rng = np.random.default_rng(0)
    src = rng.normal(size=(4096, 12, 3)).astype(np.float32)
    M = rng.normal(size=(4096, 3, 3)).astype(np.float32)
    tgt = (src @ np.swapaxes(M, -1, -2)).astype(np.float32)
    got = P.rotate_pts_batch(dev(src, cuda), dev(tgt, cuda)).cpu().numpy().astype(np.float64)
    want = PR.rotate_pts_batch(src.astype(np.float64), tgt.astype(np.float64))
    Mfull = np.swapaxes(tgt.astype(np.float64), -1, -2) @ src.astype(np.float64)
    sv = np.linalg.svd(Mfull, compute_uv=False)
    well = (sv[:, 1] - sv[:, 2]) / sv[:, 0] > 1e-3       # R is unique only when sigma2 > sigma3
    assert well.mean() > 0.9
    np.testing.assert_allclose(got[well], want[well], rtol=1e-4, atol=5e-5)
    assert np.allclose(got @ np.swapaxes(got, -1, -2), np.eye(3), atol=1e-5)
    assert np.allclose(np.linalg.det(got), 1.0, atol=1e-5)
    # zero matrix -> identity (LAPACK returns U=V=I)
    z = P.rotate_pts_batch(torch.zeros(2, 5, 3, device=cuda), torch.zeros(2, 5, 3, device=cuda))
    assert torch.equal(z.cpu(), torch.eye(3).expand(2, 3, 3))
    z2 = P.rotate_pts_2d_batch(torch.zeros(2, 5, 2, device=cuda), torch.zeros(2, 5, 2, device=cuda))
    assert torch.equal(z2.cpu(), torch.eye(2).expand(2, 2, 2))


def test_part_fit_full_size_properties(PF, cuda):
    # BASELINE cfg2/cfg3 sizes: recover a known similarity transform from noisy NOCS
    for (b, p, sym) in ((32, 1, True), (16, 2, False)):
        case = synthetic.pose_fit_case(b, p, 4096, seed=3, nocs_noise=0.002, sym=False)
        src = dev(np.swapaxes(case["nocs"], -1, -2), cuda).transpose(-1, -2)
        tgt = dev(case["cam"], cuda).transpose(1, 2).unsqueeze(1).repeat(1, p, 1, 1).transpose(-1, -2)
        model, valid = PF.part_fit_st_no_ransac(dev(case["labels"], cuda), src, tgt, dev(case["R"], cuda),
                                                {"num_parts": p, "sym": sym})
        assert valid.all()
        np.testing.assert_allclose(model["scale"].cpu().numpy(), case["s"], rtol=5e-3)
        np.testing.assert_allclose(model["translation"].cpu().numpy()[..., 0], case["t"], atol=5e-3)
        want, _ = PR.part_fit_st_no_ransac(case["labels"], case["nocs"], np.repeat(case["cam"][:, None], p, 1),
                                           case["R"], {"num_parts": p, "sym": sym})
        np.testing.assert_allclose(model["scale"].cpu().numpy(), want["scale"], **TOL)
        np.testing.assert_allclose(model["translation"].cpu().numpy(), want["translation"], **RTOL)


def test_part_fit_invalid_and_empty_parts(PF, cuda):
    case = synthetic.pose_fit_case(2, 2, 256, seed=4)
    labels = case["labels"].copy()
    labels[0][labels[0] == 1] = 2          # part 1 of cloud 0 has no points
    three = np.nonzero(labels[1] == 0)[0][3:]
    labels[1][three] = 2                   # part 0 of cloud 1 keeps exactly 3 points -> invalid (> 3 needed)
    tgt = np.repeat(case["cam"][:, None], 2, 1)
    model, valid = PF.part_fit_st_no_ransac(dev(labels, cuda), dev(case["nocs"], cuda), dev(tgt, cuda),
                                            dev(case["R"], cuda), {"num_parts": 2, "sym": False})
    want, wvalid = PR.part_fit_st_no_ransac(labels, case["nocs"], tgt, case["R"], {"num_parts": 2, "sym": False})
    assert np.array_equal(valid.cpu().numpy(), wvalid) and not wvalid[0, 1] and not wvalid[1, 0]
    np.testing.assert_allclose(model["scale"].cpu().numpy(), want["scale"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(model["translation"].cpu().numpy(), want["translation"], **RTOL)
    # NaN in the source of one part -> that part invalid, others untouched
    nocs = case["nocs"].copy()
    nocs[0, 0, np.nonzero(case["labels"][0] == 0)[0][0], 1] = np.nan
    _, v2 = PF.part_fit_st_no_ransac(dev(case["labels"], cuda), dev(nocs, cuda), dev(tgt, cuda),
                                     dev(case["R"], cuda), {"num_parts": 2, "sym": False})
    assert not v2[0, 0] and v2[1].all()
