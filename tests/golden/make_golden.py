"""Generates tests/golden/*.npz by running the REFERENCE's own Python code in this container.

Run once here (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

What is pinned:
  procrustes.npz   pose_utils/procrustes.py + pose_fit.py (torch CPU, torch.svd/LAPACK) on
                   seeded synthetic NOCS predictions: rotate_pts_batch, transform_pts_mask
                   (rotation given / None, sym / not), transform_pts_2d_mask,
                   part_fit_st_no_ransac, scale_pts_mask, translate_pts_mask.
  ops_index.npz    the reference's *live* grouping/gather (torch advanced indexing,
                   pointnet_utils.py:80-109) and its CPU three_interpolate (:46-55), which are
                   semantically equal to the CUDA kernels for given indices.
  backbone.npz     network/models/backbones.PointNet2Msg (real torch Conv/BN modules, eval mode)
                   on a reduced config, with the reference's index-producing shims
                   (farthest_point_sample, query_ball_point, three_nn -- whose CPU fallbacks are
                   NOT equal to the CUDA semantics) replaced by oracle/cpu_ref
                   so the module graph, channel orders and BN folding are pinned by reference code.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "network", "models"))
sys.path.insert(0, os.path.join(REF, "pose_utils"))
sys.path.insert(0, REF)

from captra_b200 import synthetic  # noqa: E402
from oracle import cpu_ref  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(4)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def golden_procrustes():
    import procrustes as P
    import pose_fit as PF
    out = {}
    for name, (b, p, n, sym) in {"rigid_sym": (3, 1, 512, True), "arti": (2, 2, 600, False),
                                 "arti3": (2, 3, 700, False)}.items():
        case = synthetic.pose_fit_case(b, p, n, seed=11 + p, sym=sym)
        labels = T(case["labels"])
        src = T(case["nocs"])                                   # [B,P,N,3]
        tgt = T(case["cam"]).unsqueeze(1).repeat(1, p, 1, 1)    # [B,P,N,3]
        rot = T(case["R"])
        # perturbed rotation as the tracker would supply (networks.py:227)
        model, valid = PF.part_fit_st_no_ransac(labels, src, tgt, rot, {"num_parts": p, "sym": sym})
        out[name + "/labels"] = case["labels"]
        out[name + "/source"] = case["nocs"]
        out[name + "/cam"] = case["cam"]
        out[name + "/rotation"] = case["R"]
        out[name + "/sym"] = np.array(sym)
        out[name + "/fit_scale"] = model["scale"].numpy()
        out[name + "/fit_translation"] = model["translation"].numpy()
        out[name + "/fit_valid"] = valid.numpy()
        eye = torch.cat([torch.eye(p), torch.zeros(2, p)], 0)
        mask = eye[labels].transpose(-1, -2).unsqueeze(-1)      # [B,P,N,1]
        # full Procrustes (rotation=None -> 3x3 SVD), the API surface of row a11
        R, s, t = P.transform_pts_mask(src, tgt, mask, mask, rotation=None, sym=False)
        out[name + "/full_R"], out[name + "/full_s"], out[name + "/full_t"] = R.numpy(), s.numpy(), t.numpy()
        R2, t2 = P.transform_pts_2d_mask(src[..., [0, 2]], torch.matmul(tgt, rot)[..., [0, 2]], mask)
        out[name + "/rot2d"], out[name + "/trans2d"] = R2.numpy(), t2.numpy()
        out[name + "/scale_mask"] = P.scale_pts_mask(src * mask, tgt * mask, mask).numpy()
        out[name + "/translate_mask"] = P.translate_pts_mask(src.transpose(-1, -2), tgt.transpose(-1, -2), mask).numpy()
    # unmasked batch API (row a13) and raw 3x3 / 2x2 rotations (rows a11/a12)
    rng = np.random.default_rng(5)
    s5 = rng.normal(size=(2, 2, 3, 40, 3)).astype(np.float32)
    Rg = np.stack([synthetic.random_rotation(rng) for _ in range(12)]).reshape(2, 2, 3, 3, 3).astype(np.float32)
    t5 = (0.7 * np.einsum("bphij,bphnj->bphni", Rg, s5) + rng.normal(scale=0.01, size=s5.shape) + 0.3).astype(np.float32)
    R, sc, tr = P.transform_pts_batch(T(s5), T(t5))
    out["batch/source"], out["batch/target"] = s5, t5
    out["batch/R"], out["batch/s"], out["batch/t"] = R.numpy(), sc.numpy(), tr.numpy()
    M3 = rng.normal(size=(64, 3, 3)).astype(np.float32)
    M3[0] = np.diag([1.0, 0.5, -0.25])          # reflection case
    M3[1] = np.eye(3)
    out["rot3/M"] = M3
    sc3 = T(rng.normal(size=(64, 50, 3)).astype(np.float32))
    tg3 = torch.matmul(sc3, T(M3).transpose(-1, -2))
    out["rot3/src"], out["rot3/tgt"] = sc3.numpy(), tg3.numpy()
    out["rot3/R"] = P.rotate_pts_batch(sc3, tg3).numpy()
    sc2 = T(rng.normal(size=(64, 50, 2)).astype(np.float32))
    M2 = rng.normal(size=(64, 2, 2)).astype(np.float32)
    tg2 = torch.matmul(sc2, T(M2).transpose(-1, -2))
    out["rot2/src"], out["rot2/tgt"] = sc2.numpy(), tg2.numpy()
    out["rot2/R"] = P.rotate_pts_2d_batch(sc2, tg2).numpy()
    np.savez_compressed(os.path.join(HERE, "procrustes.npz"), **out)
    print("procrustes.npz", len(out), "arrays")


def golden_ops_index():
    import pointnet_utils as PU
    assert not PU.CUDA
    rng = np.random.default_rng(3)
    B, C, N, M, K, n = 2, 5, 300, 16, 8, 40
    feats = rng.normal(size=(B, C, N)).astype(np.float32)
    gidx = rng.integers(0, N, size=(B, M, K))
    sidx = rng.integers(0, N, size=(B, M))
    out = {"feats": feats, "gidx": gidx.astype(np.int32), "sidx": sidx.astype(np.int32)}
    out["grouped"] = PU.group_operation(T(feats), T(gidx)).contiguous().numpy()
    out["gathered"] = PU.gather_operation(T(feats), T(sidx)).contiguous().numpy()
    idx3 = rng.integers(0, N, size=(B, n, 3))
    w = rng.random((B, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    out["idx3"], out["w3"] = idx3.astype(np.int32), w
    out["interp"] = PU.three_interpolate(T(feats), T(idx3), T(w)).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "ops_index.npz"), **out)
    print("ops_index.npz", len(out), "arrays")


def golden_backbone():
    import pointnet_utils as PU
    import backbones as BB

    # index-producing shims -> CUDA semantics via the oracle (see module docstring)
    PU.farthest_point_sample = lambda xyz, npoint: T(cpu_ref.furthest_point_sample(xyz.numpy(), npoint)).long()
    PU.query_ball_point = lambda radius, nsample, xyz, new_xyz: T(
        cpu_ref.ball_query(radius, nsample, xyz.contiguous().numpy(), new_xyz.contiguous().numpy())).long()

    def three_nn(xyz1, xyz2):
        d, i = cpu_ref.three_nn(xyz1.contiguous().numpy(), xyz2.contiguous().numpy())
        return T(d), T(i).long()
    PU.three_nn = three_nn

    net_cfg = {
        "sa1": {"npoint": 64, "radius_list": [0.1, 0.2, 0.4], "nsample_list": [8, 16, 32],
                "mlp_list": [[8, 8, 16], [16, 16, 24], [16, 20, 24]]},
        "sa2": {"npoint": 16, "radius_list": [0.4, 0.8], "nsample_list": [16, 32],
                "mlp_list": [[24, 24, 32], [24, 28, 32]]},
        "sa3": {"mlp": [32, 48, 64]},
        "fp3": {"mlp": [32, 32]}, "fp2": {"mlp": [32, 24]}, "fp1": {"mlp": [24, 24]},
    }
    cfg = {"pointnet": {"camera": net_cfg}, "device": "cpu"}
    out = {}
    for tag, use_xyz in (("coord", True), ("rot", False)):
        torch.manual_seed(1 if use_xyz else 2)
        net = BB.PointNet2Msg(cfg, out_dim=20, net_type="camera", use_xyz_feat=use_xyz)
        for m in net.modules():  # randomise BN running stats / affine so folding is exercised
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.normal_(0, 0.1)
        net.eval()
        pts, _ = synthetic.batch_surface_box(2, 256, seed=21)
        x = T(pts).transpose(1, 2).contiguous()  # [B,3,N]
        with torch.no_grad():
            y = net(x)
        out[tag + "/input"] = x.numpy()
        out[tag + "/output"] = y.numpy()
        for k, v in net.state_dict().items():
            out[tag + "/sd/" + k] = v.numpy()
    import json
    out["net_cfg_json"] = np.frombuffer(json.dumps(net_cfg).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "backbone.npz"), **out)
    print("backbone.npz", len(out), "arrays")


if __name__ == "__main__":
    golden_procrustes()
    golden_ops_index()
    golden_backbone()
