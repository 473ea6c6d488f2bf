"""CPU: pin oracle/frame_ref.track_step -- the checker every GPU end-to-end test and the bench's parity check
use -- against one full tracking frame run through the REFERENCE's own CoordNet and PartCanonNet
(tests/golden/frame.npz; networks.py:19-110,144-239, blocks.py:146-193) at the real widths on 4096-point clouds."""
import numpy as np
import pytest
import torch

from oracle import frame_ref

from golden_util import FEAT_STRIDE, case


@pytest.mark.parametrize("tag", ["bottle", "camera", "laptop", "bottle_t", "laptop_t"])
def test_track_step_restatement_matches_reference_networks(tag):
    cfg, trk, inp, gold = case(tag)
    sd_c = {k: v.detach() for k, v in trk.npcs_net.state_dict().items()}
    sd_r = {k: v.detach() for k, v in trk.net.state_dict().items()}
    with torch.no_grad():
        pose, inter = frame_ref.track_step(sd_c, sd_r, cfg, inp["points"], inp["points_mean"], inp["pose"])
    # both sides are torch CPU fp32 on the same weights: only the op composition differs (functional vs modules,
    # fp64 pose fit vs the reference's fp32 + LAPACK) -> tight bars
    np.testing.assert_allclose(inter["feat"][:, :, ::FEAT_STRIDE].numpy(), gold["feat_coord"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(inter["feat_rot"][:, :, ::FEAT_STRIDE].numpy(), gold["feat_rot"], rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(inter["seg"].numpy(), gold["seg"], rtol=1e-5, atol=1e-6)
    B = inp["points"].shape[0]
    np.testing.assert_allclose(inter["nocs"].reshape(B, -1, 4096).numpy(), gold["nocs"], rtol=1e-5, atol=1e-6)
    assert (inter["labels"].numpy() == gold["labels"]).all()
    np.testing.assert_allclose(pose["rotation"].numpy(), gold["pose_rotation"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pose["scale"].numpy(), gold["pose_scale"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(pose["translation"].numpy(), gold["pose_translation"], rtol=1e-4, atol=1e-5)
