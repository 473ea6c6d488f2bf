"""GPU: the per-frame glue kernels (csrc/frame_glue.cu, captra_part_fit_track) against the torch composition the
reference writes in Python (networks.py:38-46,127-141,184-232; blocks.py:181-193; pose_utils/rotations.py:300-387;
part_dof_utils.py:124-141), evaluated with the mirrors' torch functions on the same device in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _pose(B, P, gen, dev):
    from captra_b200 import synthetic
    rng = np.random.default_rng(int(torch.randint(0, 1 << 30, (1,), generator=gen)))
    R = np.stack([synthetic.random_rotation(rng) for _ in range(B * P)]).reshape(B, P, 3, 3).astype(np.float32)
    return {"rotation": torch.from_numpy(R).to(dev), "translation": torch.randn(B, P, 3, 1, generator=gen).to(dev),
            "scale": (torch.rand(B, P, generator=gen) * 0.3 + 0.2).to(dev)}


@pytest.mark.parametrize("B,P,N", [(3, 1, 4096), (2, 2, 1000), (1, 3, 77)])
def test_canonicalize(B, P, N, cuda):
    from captra_b200 import frame_ops
    from captra_b200.networks import canonicalize
    gen = torch.Generator().manual_seed(B * 100 + P)
    pts = torch.randn(B, 3, N, generator=gen).to(cuda) * 0.2
    mean = torch.randn(B, 3, 1, generator=gen).to(cuda)
    pose = _pose(B, P, gen, cuda)
    flat = {k: v.reshape((-1,) + v.shape[2:]) for k, v in pose.items()}
    rep = lambda t: t.unsqueeze(1).expand(-1, P, -1, -1).reshape((-1,) + t.shape[-2:])
    want = canonicalize(rep(pts), rep(mean), flat)                       # [B*P,3,N]
    pm, cm, dup = frame_ops.canonicalize(pts, mean, flat["rotation"], flat["translation"], flat["scale"], parts=P,
                                         want_cm=True, want_dup=True)
    torch.testing.assert_close(cm, want, rtol=1e-6, atol=1e-6)
    assert torch.equal(pm, cm.transpose(1, 2))
    assert torch.equal(dup[..., :3], pm) and torch.equal(dup[..., 3:], pm)


@pytest.mark.parametrize("B,N,nseg,P", [(3, 4096, 2, 1), (2, 1000, 2, 2), (2, 333, 5, 3)])
def test_coord_head_post(B, N, nseg, P, cuda):
    from captra_b200 import frame_ops
    gen = torch.Generator().manual_seed(N)
    seg_raw = (torch.randn(B * N, nseg, generator=gen) * 3).to(cuda)
    seg_raw[5] = seg_raw[5, 0]                    # an exact tie: the first class wins
    nocs_raw = (torch.randn(B * N, 3 * P, generator=gen) * 4).to(cuda)
    labels, nocs, seg = frame_ops.coord_head_post(seg_raw, nocs_raw, B, N)
    want_seg = torch.softmax(seg_raw.view(B, N, nseg).transpose(1, 2), dim=1)
    want_nocs = torch.sigmoid(nocs_raw.view(B, N, 3 * P).transpose(1, 2)) - 0.5
    torch.testing.assert_close(seg, want_seg, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(nocs, want_nocs, rtol=1e-6, atol=1e-7)
    assert torch.equal(labels, torch.max(seg, dim=-2)[1])                # the reference's argmax of ITS probabilities
    assert (labels != torch.max(want_seg, dim=-2)[1]).float().mean() < 1e-3
    assert labels.view(-1)[5] == 0


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("B,P,N", [(4, 1, 4096), (3, 2, 1500)])
def test_rot_head_post(sym, B, P, N, cuda):
    from captra_b200 import frame_ops
    from captra_b200.networks import RotationRegressor, convert_pred_rtvec_to_matrix
    gen = torch.Generator().manual_seed(N + P + int(sym))
    D = 3 if sym else 6
    raws = [torch.randn(B, N, D, generator=gen).to(cuda) for _ in range(P)]
    raws[0][0, 3] = 0.0                                                   # a degenerate point: the (1,0,0) backup branch
    labels = torch.randint(0, P + 1, (B, N), generator=gen).to(cuda)
    labels[B - 1] = P                                                     # last cloud: every part empty -> defaults
    rot_prev = _pose(B, P, gen, cuda)["rotation"]
    rotation, rtvec = frame_ops.rot_head_post(raws, labels, rot_prev, sym, want_rtvec=True)
    reg = RotationRegressor(128, P, symmetric=sym)
    want_vec = []
    for p in range(P):
        post = reg.post(raws[p].transpose(1, 2))                          # [B, D', N]
        mask = (labels == p).float().unsqueeze(1)
        cnt = mask.sum(-1)
        mean = (post * mask).sum(-1) / torch.clamp_min(cnt, 1.0)
        default = torch.tensor((0., 1., 0.) if sym else (1., 0., 0., 0., 1., 0., 0., 0., 1.), device=cuda)
        valid = (cnt > 0).float()
        want_vec.append(valid * mean + (1 - valid) * default.reshape(1, -1))
    want_vec = torch.stack(want_vec, dim=1)
    torch.testing.assert_close(rtvec, want_vec, rtol=1e-5, atol=1e-6)
    want_rot = torch.matmul(rot_prev, convert_pred_rtvec_to_matrix(want_vec, sym))
    torch.testing.assert_close(rotation, want_rot, rtol=1e-5, atol=2e-6)
    eye = torch.matmul(rotation.transpose(-1, -2), rotation)
    torch.testing.assert_close(eye, torch.eye(3, device=cuda).expand_as(eye), rtol=0, atol=1e-5)


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("B,P,N", [(3, 1, 4096), (2, 2, 2000)])
def test_part_fit_track_equals_fit_plus_blend(sym, B, P, N, cuda):
    from captra_b200 import frame_ops, synthetic
    from captra_b200.pose_utils.pose_fit import part_fit_st_no_ransac
    case = synthetic.pose_fit_case(B, P, N, seed=3 + P, sym=sym)
    cam = torch.from_numpy(case["cam"]).to(cuda)                          # [B,N,3]
    mean = cam.mean(1, keepdim=True).transpose(1, 2).contiguous()         # [B,3,1]
    points = (cam.transpose(1, 2) - mean).contiguous()                    # [B,3,N]
    nocs = torch.from_numpy(case["nocs"]).to(cuda).transpose(-1, -2).contiguous()   # [B,P,3,N]
    labels = torch.from_numpy(case["labels"]).to(cuda)
    labels[0] = P                                                         # cloud 0: nothing valid -> previous pose kept
    rotation = torch.from_numpy(case["R"]).to(cuda)
    gen = torch.Generator().manual_seed(1)
    prev_s = torch.rand(B, P, generator=gen).to(cuda)
    prev_t = torch.randn(B, P, 3, 1, generator=gen).to(cuda)
    scale, trans, valid = frame_ops.part_fit_track(labels, nocs, points, mean, rotation, sym, prev_s, prev_t)
    cam_points = (points + mean).unsqueeze(1).expand(-1, P, -1, -1)
    model, want_valid = part_fit_st_no_ransac(labels, nocs.transpose(-1, -2), cam_points.transpose(-1, -2), rotation,
                                              {"num_parts": P, "sym": sym})
    v = want_valid.float()
    want_s = v * model["scale"] + (1 - v) * prev_s
    v = v.unsqueeze(-1).unsqueeze(-1)
    want_t = v * model["translation"] + (1 - v) * prev_t
    assert torch.equal(valid, want_valid) and not valid[0].any()
    torch.testing.assert_close(scale, want_s, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(trans, want_t, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("N", [777, 1000])
def test_rotation_head_odd_cloud_size(N, cuda):
    """Cloud sizes that are not a multiple of the 128-row tile (the reference accepts any N)."""
    from captra_b200.networks import MLPConv1d
    torch.manual_seed(0)
    head = MLPConv1d(128, [512, 512, 256, 6]).to(cuda).eval()
    feat = torch.randn(2, 128, N, device=cuda)
    with torch.no_grad():
        want = head(feat)
        got = head.forward_pm(feat.transpose(1, 2).contiguous()).transpose(1, 2)
    torch.testing.assert_close(got, want, rtol=1e-3, atol=2e-4)


def test_tracker_odd_cloud_size(cuda):
    from captra_b200 import track
    cfg = track.make_cfg("bottle")
    trk = track.Tracker(cfg, seed=2).to(cuda).eval()
    b = track.synthetic_track_batch(2, "bottle", n=1000, seed=4)
    pose = trk.step(torch.from_numpy(b["points"]).to(cuda), torch.from_numpy(b["points_mean"]).to(cuda),
                    {k: torch.from_numpy(v).to(cuda) for k, v in b["pose"].items()})
    assert all(torch.isfinite(v).all() for v in pose.values())


@pytest.mark.parametrize("category,B", [("bottle", 5), ("laptop", 3), ("camera", 4)])
def test_track_eval_matches_reference_formulas(category, B, cuda):
    """captra_track_eval vs the reference's eval_part_full / compute_miou_loss / compute_nocs_loss formulas
    (part_dof_utils.py:40-67, metrics.py:5-47, loss.py:42-70,122-134) written out in torch."""
    from captra_b200 import frame_ops, track
    c = track.CATEGORIES[category]
    P, sym, nseg, N = c["num_parts"], c["sym"], c["num_parts"] + c["extra_dims"], 1000
    gen = torch.Generator().manual_seed(B)
    gt, pose = _pose(B, P, gen, cuda), _pose(B, P, gen, cuda)
    pose["rotation"][0] = gt["rotation"][0]                      # one exact match (acos at the clamp)
    pose["translation"][0] = gt["translation"][0] + 0.01
    seg = torch.softmax(torch.randn(B, nseg, N, generator=gen), 1).to(cuda)
    nocs = (torch.rand(B, 3 * P, N, generator=gen) - 0.5).to(cuda)
    pred_labels = torch.randint(0, nseg, (B, N), generator=gen).to(cuda)
    gt_labels = torch.randint(0, nseg, (B, N), generator=gen).to(cuda)
    gt_nocs = (torch.rand(B, 3, N, generator=gen) - 0.5).to(cuda)
    sums, per = frame_ops.track_eval(gt, pose, sym, pred={"seg": seg, "nocs": nocs, "labels": pred_labels},
                                     gt_labels=gt_labels, gt_nocs=gt_nocs, per_instance=True)
    got = frame_ops.eval_means(sums, P)
    # reference formulas
    sdiff = (gt["scale"] - pose["scale"]).abs()
    tdiff = (gt["translation"] - pose["translation"]).squeeze(-1).norm(dim=-1)
    if sym:
        d = (gt["rotation"][..., 1] * pose["rotation"][..., 1]).sum(-1)
    else:
        m = torch.matmul(gt["rotation"], pose["rotation"].transpose(-1, -2))
        d = (m[..., 0, 0] + m[..., 1, 1] + m[..., 2, 2] - 1) / 2.0
    rdiff = torch.acos(d.clamp(-1, 1)) / np.pi * 180.0
    want = {"sdiff": sdiff, "tdiff": tdiff, "rdiff": rdiff,
            "5deg5cm": ((rdiff <= 5.0) & (tdiff <= 0.05)).float(), "10deg10cm": ((rdiff <= 10.0) & (tdiff <= 0.10)).float()}
    for q, name in enumerate(("sdiff", "tdiff", "rdiff", "5deg5cm", "10deg10cm")):
        tol = dict(rtol=1e-4, atol=2e-2) if name == "rdiff" else dict(rtol=1e-5, atol=1e-6)   # acos near 1 amplifies rounding
        torch.testing.assert_close(per[..., q], want[name], **tol)
        for p in range(P):
            assert abs(got["%s_%d" % (name, p)] - float(want[name][:, p].mean())) <= tol["atol"] + tol["rtol"] * abs(float(want[name][:, p].mean()))
    onehot = torch.eye(nseg, device=cuda)[gt_labels]                              # [B,N,C]
    pr = seg.transpose(-1, -2)
    I = (pr * onehot).sum(-2)
    U = (pr + onehot).sum(-2) - I
    assert abs(got["seg_loss"] - float(1.0 - (I / (U + 1e-6)).mean())) < 1e-5
    if P > 1:
        x = nocs.transpose(-1, -2).reshape(B, N, P, 3)
        y = torch.cat([x, torch.zeros_like(x[..., :2, :])], dim=-2)
        chosen = y[torch.arange(B, device=cuda).view(-1, 1), torch.arange(N, device=cuda).view(1, -1), pred_labels]
        mask = (pred_labels < P).float()
        raw = (chosen - gt_nocs.transpose(-1, -2)).norm(dim=-1)
        want_nocs = float((raw * mask).sum() / max(float(mask.sum()), 1.0))
    else:
        want_nocs = float((nocs.transpose(-1, -2) - gt_nocs.transpose(-1, -2)).norm(dim=-1).mean())
    assert abs(got["nocs_loss"] - want_nocs) < 1e-5
    assert got["count"] == B
