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


def _install_index_shims():
    """The reference's index-producing shims -> CUDA-kernel semantics via oracle/cpu_ref (module docstring)."""
    import pointnet_utils as PU
    PU.farthest_point_sample = lambda xyz, npoint: T(cpu_ref.furthest_point_sample(xyz.contiguous().numpy(), npoint)).long()
    PU.query_ball_point = lambda radius, nsample, xyz, new_xyz: T(
        cpu_ref.ball_query(radius, nsample, xyz.contiguous().numpy(), new_xyz.contiguous().numpy())).long()

    def three_nn(xyz1, xyz2):
        d, i = cpu_ref.three_nn(xyz1.contiguous().numpy(), xyz2.contiguous().numpy())
        return T(d), T(i).long()
    PU.three_nn = three_nn


FRAME_CASES = {          # tag -> (category, clouds, weight seed, input seed, trained-like CoordNet heads)
    "bottle": ("bottle", 2, 0, 0, False),       # the bench's cfg2 tracker (seed 0) on the first two clouds of its batch 0
    "camera": ("camera", 2, 0, 40, False),      # rigid, non-symmetric: 6-D rotation head, Gram-Schmidt
    "laptop": ("laptop", 2, 0, 0, False),       # SAPIEN, two parts: per-part canonicalisation, P heads, diagonal
    # the same with track.make_trained_like applied: mixed labels, NOCS ~ canonical coordinates (well-conditioned fit)
    "bottle_t": ("bottle", 2, 3, 5, True),
    "laptop_t": ("laptop", 2, 3, 5, True),
}
FEAT_STRIDE = 16      # backbone features are stored for every 16th point (keeps the fixture small)


def golden_frame():
    """One full tracking frame (model.py:454-476) through the REFERENCE's own CoordNet and PartCanonNet
    (network/models/networks.py:19-110,144-239, blocks.py:146-193, backbones.py, pointnet_utils.py modules,
    pose_utils/{pose_fit,procrustes,part_dof_utils,rotations}.py), eval mode, at the real
    pointnet2_camera.yml widths on 4096-point clouds.  Weights: captra_b200.track.init_weights (reference
    xavier init + randomised BN running stats, keyed by state-dict name) -- NOT stored, regenerated by the
    tests and pinned by their sha256; inputs: captra_b200.track.synthetic_track_batch (regenerated too,
    digest stored).  Stored: seg, nocs, labels, canonicalised points, strided backbone features of both
    networks, per-point rotation head output, rtvec, the final pose."""
    import hashlib
    import networks as RN          # the reference's network/models/networks.py
    from captra_b200 import track
    _install_index_shims()
    out = {}
    for tag, (category, B, wseed, iseed, trained) in FRAME_CASES.items():
        cfg = track.make_cfg(category, device="cpu")
        P = cfg["num_parts"]
        npcs_net = track.init_weights(RN.CoordNet(cfg), wseed).eval()
        net = track.init_weights(RN.PartCanonNet(cfg), wseed + 1).eval()
        if trained:
            track.make_trained_like(npcs_net, P)
        batch = track.synthetic_track_batch(B, category, n=4096, seed=iseed)
        pts, mean = T(batch["points"]), T(batch["points_mean"])
        pose = {k: T(v) for k, v in batch["pose"].items()}
        root = [p for p in range(P) if cfg["obj_tree"][p] == -1][0]
        feats = {}
        h1 = npcs_net.backbone.register_forward_hook(lambda m, i, o: feats.__setitem__("coord", o.detach()))
        h2 = net.regress_net.encoder.register_forward_hook(lambda m, i, o: feats.__setitem__("rot", o.detach()))
        with torch.no_grad():
            # model.py:454-476 (EvalTrackModel.forward loop body), npcs_net then net
            canon = {k: pose[k][:, root] for k in ("rotation", "translation", "scale")}
            pred = npcs_net({"points": pts, "points_mean": mean, "canon_pose": canon})
            pred_labels = torch.max(pred["seg"], dim=-2)[1]
            pred_npcs = pred["nocs"].reshape(B, P, 3, -1)
            res = net({"points": pts, "points_mean": mean, "state": {"part": pose}, "pred_labels": pred_labels,
                       "pred_nocs": pred_npcs}, test_mode=True)
        h1.remove()
        h2.remove()
        out[tag + "/meta"] = np.array([B, wseed, iseed, int(trained)])
        out[tag + "/input_digest"] = np.frombuffer(hashlib.sha256(batch["points"].tobytes() + batch["pose"]["rotation"].tobytes()).digest(), np.uint8)
        out[tag + "/coord_sd_digest"] = np.frombuffer(bytes.fromhex(track.state_dict_digest(npcs_net)), np.uint8)
        out[tag + "/rot_sd_digest"] = np.frombuffer(bytes.fromhex(track.state_dict_digest(net)), np.uint8)
        out[tag + "/canon_points"] = pred["points"].numpy()
        out[tag + "/seg"] = pred["seg"].numpy()
        out[tag + "/nocs"] = pred["nocs"].numpy()
        out[tag + "/labels"] = pred_labels.numpy().astype(np.int16)
        out[tag + "/feat_coord"] = feats["coord"][:, :, ::FEAT_STRIDE].contiguous().numpy()
        out[tag + "/feat_rot"] = feats["rot"][:, :, ::FEAT_STRIDE].contiguous().numpy()
        out[tag + "/point_rotation"] = res["point_rotation"][..., ::FEAT_STRIDE, :, :].contiguous().numpy()
        for k in ("rotation", "scale", "translation"):
            out[tag + "/pose_" + k] = res["part"][k].numpy()
        print(tag, "labels hist", np.bincount(pred_labels.numpy().ravel()), "scale", res["part"]["scale"].numpy().ravel())
    np.savez_compressed(os.path.join(HERE, "frame.npz"), **out)
    print("frame.npz", len(out), "arrays", os.path.getsize(os.path.join(HERE, "frame.npz")) // 1024, "KiB")


sys.path.insert(0, os.path.dirname(HERE))
from golden_util import CROP_CASES  # noqa: E402


def golden_crop():
    """datasets/nocs_data/nocs_data_process.py:148-164 crop_ball_from_depth_image (-> crop_ball_from_pts :92-109,
    get_proj_corners :129-143, nocs_utils.backproject, data_utils.farthest_point_sample :138-158) -- the REFERENCE's own
    functions, run here on synthetic depth frames (captra_b200.synthetic.depth_scene).  Their CUDA branch is taken
    (torch.cuda.is_available patched to True) with the FPS kernel call replaced by oracle/cpu_ref (the kernel's
    semantics); the permutation numpy's global RNG hands out is recorded so the device path can be fed the same one.
    trimesh / matplotlib / pylab (not installed; imported by data_utils.py for unrelated code) are stubbed."""
    import importlib.abc
    import importlib.machinery
    import types
    from unittest import mock

    class _Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return mock.MagicMock()

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        P = ("trimesh", "matplotlib", "mpl_toolkits", "pylab")

        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in self.P:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)

        def create_module(self, spec):
            m = _Stub(spec.name)
            m.__path__ = []
            return m

        def exec_module(self, module):
            pass
    sys.meta_path.insert(0, Finder())
    sys.path[:0] = [os.path.join(REF, "datasets", "nocs_data"), os.path.join(REF, "datasets")]
    import nocs_data_process as NDP
    import data_utils as DU
    DU.farthest_point_sample_cuda = lambda xyz, npoint: T(cpu_ref.furthest_point_sample(np.ascontiguousarray(xyz.numpy()), npoint)).long()
    out = {}
    for name, (scene, off, radius, num_points) in CROP_CASES.items():
        depth, mask, c, K = synthetic.depth_scene(**scene)
        center = c + np.array(off)
        perms = []
        real_perm = np.random.permutation

        def rec_perm(n):
            p = real_perm(n)
            perms.append(p)
            return p
        np.random.seed(100 + scene["seed"])
        with mock.patch.object(torch.cuda, "is_available", return_value=True), mock.patch.object(np.random, "permutation", rec_perm):
            pts, obj_mask = NDP.crop_ball_from_depth_image(depth, mask, center.copy(), radius, cam_intrinsics=K,
                                                           num_points=num_points, device="cpu")
        assert len(perms) <= 1
        out[name + "/pts"] = pts
        out[name + "/obj_mask"] = obj_mask.astype(np.int32)
        out[name + "/center"] = center
        if perms:
            out[name + "/perm_len"] = np.array(len(perms[0]))
            out[name + "/perm_head"] = perms[0][:5 * num_points].astype(np.int32)
        print(name, "points", pts.shape, "object fraction %.2f" % obj_mask.mean(), "perm" if perms else "")
    np.savez_compressed(os.path.join(HERE, "crop.npz"), **out)
    print("crop.npz", len(out), "arrays", os.path.getsize(os.path.join(HERE, "crop.npz")) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["procrustes", "ops_index", "backbone", "frame", "crop"]
    for name in which:
        globals()["golden_" + name]()
