"""CPU: pin oracle/crop_ref.py (numpy restatement of the reference's per-frame crop) against the outputs of the
reference's OWN crop_ball_from_depth_image / backproject / farthest_point_sample (tests/golden/crop.npz)."""
import os

import numpy as np
import pytest

from captra_b200 import synthetic
from oracle import crop_ref

from golden_util import CROP_CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop.npz")


def full_perm(g, name, num_points):
    """The reference's recorded permutation: only its first 5 * num_points entries are used (data_utils.py:147)."""
    if name + "/perm_head" not in g:
        return None
    return g[name + "/perm_head"].astype(np.int64)


@pytest.mark.parametrize("name", sorted(CROP_CASES))
def test_crop_restatement_matches_reference(name):
    g = np.load(GOLD)
    scene, off, radius, num_points = CROP_CASES[name]
    depth, mask, c, K = synthetic.depth_scene(**scene)
    center = c + np.array(off)
    np.testing.assert_array_equal(center, g[name + "/center"])
    pts, obj_mask, _ = crop_ref.crop_ball_from_depth_image(depth, mask, center, radius, K, num_points, perm=full_perm(g, name, num_points))
    np.testing.assert_array_equal(pts, g[name + "/pts"])          # same numpy expressions, same order: bit-exact
    np.testing.assert_array_equal(obj_mask, g[name + "/obj_mask"])
