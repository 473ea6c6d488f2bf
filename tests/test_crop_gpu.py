"""GPU: the device-side crop (csrc/crop.cu, captra_b200/data_crop.py) against the outputs of the reference's own
crop_ball_from_depth_image on the same depth frames and the same random permutation (tests/golden/crop.npz), and
against the CPU restatement oracle/crop_ref.py."""
import os

import numpy as np
import pytest
import torch

from captra_b200 import synthetic

from golden_util import CROP_CASES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


@pytest.mark.parametrize("name", sorted(CROP_CASES))
def test_device_crop_matches_reference(name, cuda):
    from captra_b200 import data_crop
    g = np.load(GOLD)
    scene, off, radius, num_points = CROP_CASES[name]
    depth, mask, c, K = synthetic.depth_scene(**scene)
    center = c + np.array(off)
    perm = torch.from_numpy(g[name + "/perm_head"].astype(np.int64)).to(cuda) if name + "/perm_head" in g else None
    pts, obj_mask, info = data_crop.crop_ball_from_depth_image(torch.from_numpy(depth).to(cuda), torch.from_numpy(mask).to(cuda), center, radius,
                                                               cam_intrinsics=K, num_points=num_points, perm=perm, return_info=True)
    want = g[name + "/pts"]
    assert pts.shape == want.shape and pts.dtype == torch.float64
    # the back-projection is the same fp64 expression; numpy's matmul may fuse / order its three products differently:
    # 1e-12 relative.  The SELECTION (which pixels, in which order, which FPS picks) must be identical.
    np.testing.assert_allclose(pts.cpu().numpy(), want, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(obj_mask.cpu().numpy(), g[name + "/obj_mask"])
    assert info["n"] > 0


def test_device_crop_recurses_when_window_is_empty(cuda):
    """nocs_data_process.py:159-160: no point at all -> retry with radius * 1.2 (here: depth holes over the whole first window)."""
    from captra_b200 import data_crop
    from oracle import crop_ref
    depth, mask, c, K = synthetic.depth_scene(seed=9, obj_radius=0.1, obj_depth=0.9, holes=0.0)
    win = data_crop.get_proj_corners(depth.shape, c, 0.05, K)
    depth[win[0, 0]:win[1, 0] + 1, win[0, 1]:win[1, 1] + 1] = 0.0
    perm = np.random.default_rng(0).permutation(1 << 16)          # more than any crop here; only its head is used
    want_pts, want_mask, _ = crop_ref.crop_ball_from_depth_image(depth, mask, c, 0.05, K, 256, perm=perm[perm < 2340])
    pts, obj_mask, info = data_crop.crop_ball_from_depth_image(torch.from_numpy(depth).to(cuda), torch.from_numpy(mask).to(cuda), c, 0.05,
                                                              cam_intrinsics=K, num_points=256, perm=torch.from_numpy(perm[perm < 2340]).to(cuda),
                                                              return_info=True)
    assert info["n"] == 2340
    np.testing.assert_allclose(pts.cpu().numpy(), want_pts, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(obj_mask.cpu().numpy(), want_mask)
