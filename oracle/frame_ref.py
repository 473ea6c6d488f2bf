"""CPU restatement of one tracking frame in the REFERENCE's own structure (unfused, channel-major):
network/models/pointnet_utils.py:191-343 (SA-MSG / FP / group-all with separate conv, BN, ReLU,
max), backbones.py:55-69, networks.py:34-46 (CoordNet), :123-141 + blocks.py:181-193 (RotNet heads,
all P heads on all B*P copies, diagonal kept: networks.py:200-203), :211-232 (compose + pose fit).

TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.c).  Index-producing ops use oracle/cpu_ref (the CUDA
kernels' semantics, which is what the tracker runs; the reference's torch CPU fallbacks are not
equal to them); dense math is torch CPU fp32 functional ops on the weights of
a reference-keyed state_dict; the pose fit is oracle/pose_ref (fp64).  Used by tests as the
end-to-end checker and by bench.py as the timed `cpu_baseline` / `--impl reference` arm ("port").
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cpu_ref, pose_ref

BN_EPS = 1e-5


def _T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _conv_bn_relu(x, sd, conv, bn, two_d):
    w, b = sd[conv + ".weight"], sd[conv + ".bias"]
    x = F.conv2d(x, w, b) if two_d else F.conv1d(x, w, b)
    x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.0, BN_EPS)
    return F.relu(x)


def _index_points(points, idx):
    """pointnet_utils.py:82-97: points [B,N,C], idx [B,...] -> [B,...,C]."""
    B = points.shape[0]
    view = [B] + [1] * (idx.dim() - 1)
    batch = torch.arange(B).view(view).expand_as(idx)
    return points[batch, idx, :]


def sa_msg(sd, pre, cfg, xyz, points):
    """pointnet_utils.py:213-250.  xyz [B,3,N], points [B,D,N] (D may be 0)."""
    B, C, N = xyz.shape
    S = cfg["npoint"]
    xyz_t = xyz.permute(0, 2, 1).contiguous()
    fps_idx = _T(cpu_ref.furthest_point_sample(xyz_t.numpy(), S)).long()
    new_xyz_t = _index_points(xyz_t, fps_idx)                       # [B,S,3]
    new_xyz = new_xyz_t.permute(0, 2, 1)
    outs = []
    for i, radius in enumerate(cfg["radius_list"]):
        K = cfg["nsample_list"][i]
        gidx = _T(cpu_ref.ball_query(radius, K, xyz_t.numpy(), new_xyz_t.contiguous().numpy())).long()
        grouped_xyz = _index_points(xyz_t, gidx).permute(0, 3, 1, 2) - new_xyz.reshape(B, C, S, 1)
        if points is not None:
            g = torch.cat([_index_points(points.permute(0, 2, 1), gidx).permute(0, 3, 1, 2), grouped_xyz], dim=1)
        else:
            g = grouped_xyz
        for j in range(len(cfg["mlp_list"][i])):
            g = _conv_bn_relu(g, sd, "%s.conv_blocks.%d.%d" % (pre, i, j), "%s.bn_blocks.%d.%d" % (pre, i, j), True)
        outs.append(torch.max(g, -1)[0])
    return new_xyz, torch.cat(outs, dim=1)


def sa_all(sd, pre, nlayers, xyz, points):
    """pointnet_utils.py:319-343: channels [xyz, points], max over all points."""
    x = torch.cat([xyz, points], dim=1).unsqueeze(-1)                # [B,3+D,N,1]
    for i in range(nlayers):
        x = _conv_bn_relu(x, sd, "%s.mlp_convs.%d" % (pre, i), "%s.mlp_bns.%d" % (pre, i), True)
    return torch.zeros(xyz.shape[0], 3, 1), torch.max(x, 2)[0]


def fp(sd, pre, nlayers, xyz1, xyz2, points1, points2):
    """pointnet_utils.py:265-299."""
    N, S = xyz1.shape[2], xyz2.shape[2]
    if S == 1:
        interp = points2.repeat(1, 1, N)
    else:
        x1 = xyz1.permute(0, 2, 1).contiguous().numpy()
        x2 = xyz2.permute(0, 2, 1).contiguous().numpy()
        dist, idx = cpu_ref.three_nn(x1, x2)
        w = cpu_ref.interp_weights(dist)
        interp = _T(cpu_ref.three_interpolate(points2.contiguous().numpy(), idx, w))
    x = torch.cat([points1, interp], dim=-2) if points1 is not None else interp
    for i in range(nlayers):
        x = _conv_bn_relu(x, sd, "%s.mlp_convs.%d" % (pre, i), "%s.mlp_bns.%d" % (pre, i), False)
    return x


def backbone(sd, pre, net_cfg, x, use_xyz_feat):
    """backbones.py:55-69.  x [B,3,N] -> [B,out_dim,N]."""
    p = (pre + ".") if pre else ""
    l0_xyz = x
    l0_points = x if use_xyz_feat else x[:, 3:]
    l1_xyz, l1_points = sa_msg(sd, p + "sa1", net_cfg["sa1"], l0_xyz, l0_points)
    l2_xyz, l2_points = sa_msg(sd, p + "sa2", net_cfg["sa2"], l1_xyz, l1_points)
    l3_xyz, l3_points = sa_all(sd, p + "sa3", len(net_cfg["sa3"]["mlp"]), l2_xyz, l2_points)
    l2_points = fp(sd, p + "fp3", len(net_cfg["fp3"]["mlp"]), l2_xyz, l3_xyz, l2_points, l3_points)
    l1_points = fp(sd, p + "fp2", len(net_cfg["fp2"]["mlp"]), l1_xyz, l2_xyz, l1_points, l2_points)
    l0_points = fp(sd, p + "fp1", len(net_cfg["fp1"]["mlp"]), l0_xyz, l1_xyz, torch.cat([l0_xyz, l0_points], 1), l1_points)
    return _conv_bn_relu(l0_points, sd, p + "conv1", p + "bn1", False)


def _canonicalize(cam, mean, pose):
    cam = cam + mean - pose["translation"]
    cam = torch.matmul(pose["rotation"].transpose(-1, -2), cam)
    return cam / pose["scale"].unsqueeze(-1).unsqueeze(-1)


def _normalize(v):
    mag = torch.norm(v, p=2, dim=1, keepdim=True)
    valid = (mag > 1e-8).float()
    backup = torch.tensor([1.0, 0.0, 0.0]).view(1, 3).expand_as(v)
    return v / torch.clamp(mag, min=1e-8) * valid + backup * (1 - valid)


def _rot_from_3d(vec):
    """rotations.py:375-387."""
    y = _normalize(vec)
    x_raw = torch.zeros_like(y)
    x_raw[:, 0] = 1.0
    z = _normalize(torch.cross(x_raw, y, dim=1))
    x = torch.cross(y, z, dim=1)
    return torch.stack((x, y, z), dim=2)


def _rot_from_6d(p):
    """rotations.py:330-343."""
    x = _normalize(p[:, 0:3])
    z = _normalize(torch.cross(x, p[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


def _rot_from_matrix(m):
    """rotations.py:354-372."""
    def proj(u, a):
        return ((u * a).sum(1) / torch.clamp((u * u).sum(1), min=1e-8)).unsqueeze(1) * u
    a1, a2, a3 = m[:, :, 0], m[:, :, 1], m[:, :, 2]
    u2 = a2 - proj(a1, a2)
    u3 = a3 - proj(a1, a3) - proj(u2, a3)
    return torch.stack((_normalize(a1), _normalize(u2), _normalize(u3)), dim=2)


def rot_head(sd, pre, x):
    """blocks.py:146-165 (MLPConv1d, gn=True): conv, GroupNorm(C/2), ReLU x3, conv."""
    for li, has_norm in ((0, True), (3, True), (6, True), (9, False)):
        x = F.conv1d(x, sd["%s.model.%d.weight" % (pre, li)], sd["%s.model.%d.bias" % (pre, li)])
        if has_norm:
            g = "%s.model.%d" % (pre, li + 1)
            x = F.relu(F.group_norm(x, x.shape[1] // 2, sd[g + ".weight"], sd[g + ".bias"], 1e-5))
    return x


def track_step(sd_coord, sd_rot, cfg, points, points_mean, last_pose):
    """One iteration of model.py:409-478 for a batch.  sd_coord / sd_rot: state dicts of CoordNet /
    PartCanonNet (reference key names), tensors on CPU.  points [B,3,N], points_mean [B,3,1],
    last_pose dict of CPU tensors.  Returns (new pose dict, intermediates dict)."""
    P, sym = cfg["num_parts"], cfg["obj_sym"]
    net_cfg = cfg["pointnet"]["camera"]
    root = [p for p in range(P) if cfg["obj_tree"][p] == -1][0]
    B = points.shape[0]
    # --- CoordNet (networks.py:34-46) ---
    canon = {k: last_pose[k][:, root] for k in ("rotation", "translation", "scale")}
    cam = _canonicalize(points, points_mean, canon)
    feat = backbone(sd_coord, "backbone", net_cfg, cam, True)
    seg = F.softmax(F.conv1d(feat, sd_coord["seg_head.0.weight"], sd_coord["seg_head.0.bias"]), dim=1)
    h = _conv_bn_relu(feat, sd_coord, "nocs_head.0", "nocs_head.1", False)
    nocs = torch.sigmoid(F.conv1d(h, sd_coord["nocs_head.3.weight"], sd_coord["nocs_head.3.bias"])) - 0.5
    pred_labels = torch.max(seg, dim=-2)[1]
    pred_npcs = nocs.reshape(B, P, 3, -1)
    # --- PartCanonNet (networks.py:156-232) ---
    part_pose = last_pose
    canon_p = {k: part_pose[k].reshape((-1,) + part_pose[k].shape[2:]) for k in ("rotation", "translation", "scale")}
    cam_rep = points.unsqueeze(1).repeat(1, P, 1, 1).reshape(-1, 3, points.shape[-1])
    mean_rep = points_mean.unsqueeze(1).repeat(1, P, 1, 1).reshape(-1, 3, 1)
    lab_rep = pred_labels.unsqueeze(1).repeat(1, P, 1).reshape(-1, pred_labels.shape[-1])
    cam_rep = _canonicalize(cam_rep, mean_rep, canon_p)
    feat_r = backbone(sd_rot, "regress_net.encoder", net_cfg, cam_rep, False)
    raw = torch.stack([rot_head(sd_rot, "regress_net.pose_pred.rtvec_head.%d" % p, feat_r) for p in range(P)], dim=1)
    rt = raw.transpose(-1, -2)                                       # [B*P, P, N, D]
    shape = rt.shape
    if sym:
        rt = _normalize(rt.reshape(-1, 3)).reshape(shape).transpose(-1, -2)
    else:
        rt = _rot_from_6d(rt.reshape(-1, 6)).reshape(shape[:-1] + (-1,)).transpose(-1, -2)
    eye = torch.cat([torch.eye(P), torch.zeros(2, P)], dim=0)
    part_mask = eye[lab_rep].transpose(-1, -2).unsqueeze(-2)          # [B*P, P, 1, N]
    valid_mask = (part_mask.sum(dim=(-1, -2)) > 0).float().unsqueeze(-1)
    weighted = (rt * part_mask).sum(-1) / torch.clamp_min(part_mask.sum(-1), 1.0)
    default = torch.tensor((0., 1., 0.)) if sym else torch.eye(3).reshape(-1)
    weighted = valid_mask * weighted + (1.0 - valid_mask) * default.reshape(1, 1, -1)
    if sym:
        rot = _rot_from_3d(weighted.reshape(-1, 3)).reshape(weighted.shape[:-1] + (3, 3))
    else:
        rot = _rot_from_matrix(weighted.reshape(-1, 3, 3)).reshape(weighted.shape[:-1] + (3, 3))
    rot = rot.reshape(B, P, P, 3, 3)
    ar = torch.arange(P)
    delta = rot[:, ar, ar]                                             # diagonal (networks.py:200-203)
    rotation = torch.matmul(part_pose["rotation"], delta)
    cam_points = (points + points_mean).unsqueeze(1).repeat(1, P, 1, 1)
    model, valid = pose_ref.part_fit_st_no_ransac(
        pred_labels.numpy(), pred_npcs.transpose(-1, -2).numpy(), cam_points.transpose(-1, -2).numpy(),
        rotation.numpy().astype(np.float64), {"num_parts": P, "sym": sym})
    v = _T(valid.astype(np.float32))
    scale = v * _T(model["scale"].astype(np.float32)) + (1 - v) * part_pose["scale"]
    v3 = v.unsqueeze(-1).unsqueeze(-1)
    trans = v3 * _T(model["translation"].astype(np.float32)) + (1 - v3) * part_pose["translation"]
    pose = {"rotation": rotation, "scale": scale, "translation": trans}
    return pose, {"labels": pred_labels, "nocs": pred_npcs, "seg": seg, "feat": feat, "feat_rot": feat_r, "valid": valid}
