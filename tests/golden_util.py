"""Shared loader for tests/golden/frame.npz (made by tests/golden/make_golden.py::golden_frame by running the
REFERENCE's own CoordNet / PartCanonNet).  Weights and inputs are regenerated from seeds and checked against the
digests stored with the golden, so a mismatch in the generators fails loudly instead of as a numeric diff."""
import hashlib
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame.npz")
FEAT_STRIDE = 16
CATEGORY = {"bottle": "bottle", "camera": "camera", "laptop": "laptop", "bottle_t": "bottle", "laptop_t": "laptop"}
_cache = {}

CROP_CASES = {   # name -> (scene kwargs, centre offset from the sphere centre [m], crop radius [m], num_points)
    "large_thinned": (dict(seed=1, obj_radius=0.25, obj_depth=0.6), (0.0, 0.0, 0.1), 0.3, 4096),      # > 5 * 4096 points: random subset, then FPS
    "medium": (dict(seed=2, obj_radius=0.12, obj_depth=0.8), (0.0, 0.0, 0.05), 0.15, 4096),          # 4096 < n <= 20480: plain FPS
    "small_tiled": (dict(seed=3, obj_radius=0.05, obj_depth=1.2), (0.0, 0.0, 0.02), 0.06, 4096),      # n < 4096: idx doubled until >= 4096
    "radius_grows": (dict(seed=4, obj_radius=0.1, obj_depth=0.9), (0.0, 0.0, 0.165), 0.05, 1024),     # < 10 points at first: radius *= 1.10
    "take_all": (dict(seed=5, obj_radius=0.1, obj_depth=0.9), (0.0, 0.0, 0.45), 0.05, 512),           # still empty after ten tries: every point of the window
    "no_resample": (dict(seed=6, obj_radius=0.12, obj_depth=0.8), (0.0, 0.0, 0.05), 0.1, None),       # num_points=None: the ball, no FPS
}


def load():
    if "g" not in _cache:
        _cache["g"] = dict(np.load(GOLD))
    return _cache["g"]


def case(tag, device="cpu"):
    """-> (cfg, tracker (mirror modules with the golden's weights, eval), inputs dict of CPU tensors, golden dict)."""
    from captra_b200 import track
    g = load()
    B, wseed, iseed, trained = (int(v) for v in g[tag + "/meta"])
    category = CATEGORY[tag]
    cfg = track.make_cfg(category, device=str(device))
    trk = track.Tracker(cfg, seed=wseed)
    if trained:
        track.make_trained_like(trk.npcs_net, cfg["num_parts"])
    assert bytes.fromhex(track.state_dict_digest(trk.npcs_net)) == g[tag + "/coord_sd_digest"].tobytes(), "CoordNet weights differ from the golden's"
    assert bytes.fromhex(track.state_dict_digest(trk.net)) == g[tag + "/rot_sd_digest"].tobytes(), "PartCanonNet weights differ from the golden's"
    batch = track.synthetic_track_batch(B, category, n=4096, seed=iseed)
    dg = hashlib.sha256(batch["points"].tobytes() + batch["pose"]["rotation"].tobytes()).digest()
    assert dg == g[tag + "/input_digest"].tobytes(), "synthetic inputs differ from the golden's"
    inputs = {"points": torch.from_numpy(batch["points"]), "points_mean": torch.from_numpy(batch["points_mean"]),
              "pose": {k: torch.from_numpy(v) for k, v in batch["pose"].items()}}
    gold = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + "/")}
    return cfg, trk.eval(), inputs, gold


def pose_tolerance(gold, inputs, cfg, nocs_got):
    """Absolute tolerances (tol_scale [B,P], tol_translation [B,P]) for a pose computed from NOCS predictions
    `nocs_got` that differ slightly from the golden's.  Derivation (pose_utils/procrustes.py:117-120,123-129,158-162):
    with the centred masked sets a_i = R s_i (source = predicted NOCS) and b_i (target = camera points),
        scale = sum a_i.b_i / (sum |a_i|^2 + eps),      translation = mean(b) - scale R mean(s).
    A perturbation d_i of the source moves the scale, to first order, by
        d scale = sum d_i.(b_i - 2 scale a_i) / sum |a_i|^2,   so by Cauchy-Schwarz
        |d scale| <= ||d||_2 (||b||_2 + 2 |scale| ||a||_2) / ||a||_2^2
    (for symmetric categories the in-plane refinement angle maximises sum a.b, so it enters at second order only),
    and the translation by at most |mean s| |d scale| + |scale| |mean d|.  ||d||_2 is MEASURED from nocs_got against
    the golden NOCS on the part's own points.  On top of that sits the north star's 1e-4 relative bar for the fit
    itself (which tests/test_pose_gpu.py checks on identical inputs).  With NOCS that follow the geometry (the
    *_t cases, like a trained network) the amplification is O(1) and the bound stays ~1e-4; with raw random
    weights the NOCS are uncorrelated with the cloud, ||b||/||a|| is large, and a relative bar on the scale is not
    meaningful -- this bound is what fp32 arithmetic can promise there."""
    P = cfg["num_parts"]
    labels = gold["labels"].astype(np.int64)                                   # [B,N]
    B, N = labels.shape
    nocs = gold["nocs"].astype(np.float64).reshape(B, P, 3, N)
    got = np.asarray(nocs_got, dtype=np.float64).reshape(B, P, 3, N)
    cam = (inputs["points"] + inputs["points_mean"]).numpy().astype(np.float64)  # [B,3,N]
    tol_s, tol_t = np.zeros((B, P)), np.zeros((B, P))
    for b in range(B):
        for p in range(P):
            m = labels[b] == p
            s = abs(float(gold["pose_scale"][b, p]))
            t = float(np.linalg.norm(gold["pose_translation"][b, p]))
            tol_s[b, p], tol_t[b, p] = 1e-4 * s + 1e-6, 1e-4 * max(t, 1.0)
            if m.sum() <= 3:
                continue
            a, d, tg = nocs[b, p][:, m], (got[b, p] - nocs[b, p])[:, m], cam[b][:, m]
            ac, tc = a - a.mean(1, keepdims=True), tg - tg.mean(1, keepdims=True)
            dc = d - d.mean(1, keepdims=True)
            na, nb, nd = np.linalg.norm(ac), np.linalg.norm(tc), np.linalg.norm(dc)
            ds = 1.5 * nd * (nb + 2 * s * na) / (na * na + 1e-6)
            tol_s[b, p] += ds
            tol_t[b, p] += np.linalg.norm(a.mean(1)) * (tol_s[b, p]) + s * np.linalg.norm(d.mean(1))
    return tol_s, tol_t
