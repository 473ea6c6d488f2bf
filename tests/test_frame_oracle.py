"""CPU: pin oracle/frame_ref.py's restatement of the backbone against the output of the
REFERENCE's own PointNet2Msg modules stored in tests/golden/backbone.npz."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import frame_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "backbone.npz")


@pytest.mark.parametrize("tag", ["coord", "rot"])
def test_backbone_restatement_matches_reference_modules(tag):
    g = np.load(GOLD)
    net_cfg = json.loads(bytes(g["net_cfg_json"]).decode())
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/sd/")}
    x = torch.from_numpy(g[tag + "/input"])
    with torch.no_grad():
        y = frame_ref.backbone(sd, "", net_cfg, x, tag == "coord")
    np.testing.assert_allclose(y.numpy(), g[tag + "/output"], rtol=1e-5, atol=1e-6)
