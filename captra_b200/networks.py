"""Mirror of the tracking-time part of network/models/networks.py + blocks.py: CoordNet,
RotationRegressionBackbone / RotationRegressor and PartCanonNet with the same constructor
arguments (cfg dict), the same forward(dict) contract and the same state-dict keys, so the
reference's checkpoints load and EvalTrackModel.forward (model.py:386-478) can drive them.

Scope note: in eval mode the per-point heads run on this package's
kernels too -- seg / NOCS heads as BN-folded fused MLPs, the RotationRegressor heads as one tcgen05
GEMM per conv1d with GroupNorm + ReLU folded into the consumer's operand load
(captra_group_norm_affine + captra_point_mlp_affine).  With CAPTRA_MLP_IMPL=0, in training mode or
under autograd the torch modules run, exactly as in the reference.  One exact saving is taken: the
reference evaluates all P rotation heads on all B*P canonicalised clouds and keeps the diagonal
(networks.py:200-203); head p is evaluated only on part p's copy here, which yields the same tensors.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, frame_ops
from . import mlp as _mlp
from .backbones import PointNet2Msg
from .mlp import PackedMLP, fold_conv_bn, group_norm_affine, group_norm_finalize
from .pointnet_utils import _FusedCache, _needs_autograd
from .pose_utils.pose_fit import part_fit_st_no_ransac
from .pose_utils.procrustes import rot_around_yaxis_to_3d, scale_pts_mask, transform_pts_2d_mask, translate_pts_mask


# ---- pose_utils/rotations.py:300-387, part_dof_utils.py:124-141 (elementwise glue) ---------------
def normalize_vector(v):
    """rotations.py:300-312."""
    mag = torch.norm(v, p=2, dim=1, keepdim=True)
    valid = (mag > 1e-8).float()
    backup = torch.zeros_like(v)          # (1, 0, 0), built on the device (CUDA-graph capturable)
    backup[:, 0] = 1.0
    return (v / torch.clamp(mag, min=1e-8)) * valid + backup * (1 - valid)


def _proj(u, a):
    """rotations.py:344-351."""
    top = (u * a).sum(1)
    bottom = torch.clamp((u * u).sum(1), min=1e-8)
    return (top / bottom).unsqueeze(1) * u


def compute_rotation_matrix_from_ortho6d(poses):
    """rotations.py:330-343."""
    x = normalize_vector(poses[:, 0:3])
    z = normalize_vector(torch.cross(x, poses[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


def compute_rotation_matrix_from_matrix(m):
    """rotations.py:354-372: Gram-Schmidt on the columns."""
    a1, a2, a3 = m[:, :, 0], m[:, :, 1], m[:, :, 2]
    u2 = a2 - _proj(a1, a2)
    u3 = a3 - _proj(a1, a3) - _proj(u2, a3)
    return torch.stack((normalize_vector(a1), normalize_vector(u2), normalize_vector(u3)), dim=2)


def compute_rotation_matrix_from_3d(vec):
    """rotations.py:375-387: y = v/|v|, z = x_raw x y, x = y x z."""
    y = normalize_vector(vec)
    x_raw = torch.zeros_like(y)
    x_raw[..., 0] = 1.0
    z = normalize_vector(torch.cross(x_raw, y, dim=1))
    x = torch.cross(y, z, dim=1)
    return torch.stack((x, y, z), dim=2)


def convert_pred_rtvec_to_matrix(pred, sym):
    """part_dof_utils.py:137-141."""
    if sym:
        return compute_rotation_matrix_from_3d(pred.reshape(-1, pred.shape[-1])).reshape(pred.shape[:-1] + (3, 3))
    return compute_rotation_matrix_from_matrix(pred.reshape(-1, 3, 3)).reshape(pred.shape[:-1] + (3, 3))


def canonicalize(cam, points_mean, pose):
    """networks.py:38-41 / :184-187: (cam + mean - t) -> R^T . -> / s.   cam [B,3,N]."""
    cam = cam + points_mean - pose['translation']
    cam = torch.matmul(pose['rotation'].transpose(-1, -2), cam)
    return cam / pose['scale'].unsqueeze(-1).unsqueeze(-1)


# ---- blocks.py:118-193 -----------------------------------------------------------------------
def get_point_mlp(in_dim, out_dim, dims, acti='none'):
    """blocks.py:118-135 with dropout=None (networks.py:29-32): conv1d(+bn+relu)* + conv1d + acti."""
    layers, dims = [], [in_dim] + list(dims) + [out_dim]
    for i in range(len(dims) - 2):
        layers += [nn.Conv1d(dims[i], dims[i + 1], 1), nn.BatchNorm1d(dims[i + 1]), nn.ReLU(inplace=True)]
    layers.append(nn.Conv1d(dims[-2], dims[-1], 1))
    if acti == 'sigmoid':
        layers.append(nn.Sigmoid())
    return nn.Sequential(*layers)


class MLPConv1d(nn.Module):
    """blocks.py:146-165 with gn=True: conv1d + GroupNorm(C/2 groups) + ReLU per hidden layer."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        layers, last = [], in_channel
        for i, out_channel in enumerate(mlp):
            layers.append(nn.Conv1d(last, out_channel, 1))
            if i != len(mlp) - 1:
                layers += [nn.GroupNorm(out_channel // 2, out_channel), nn.ReLU(inplace=True)]
            last = out_channel
        self.model = nn.Sequential(*layers)
        self.out_channel = last

    def forward(self, x):
        return self.model(x)

    def forward_pm(self, feat_pm):
        """Fused inference path on point-major features [B,N,C] -> raw head output [B,N,D].
        Every conv1d is one tcgen05 launch; GroupNorm + ReLU never materialise: the statistics of a
        layer's pre-norm output become a per-(cloud, channel) affine that the next layer applies
        while it loads its operand."""
        B, N, C = feat_pm.shape
        convs = [m for m in self.model if isinstance(m, nn.Conv1d)]
        gns = [m for m in self.model if isinstance(m, nn.GroupNorm)]
        if not hasattr(self, "_cache"):
            self._cache = _FusedCache()
        # GroupNorm-on-load exists on the tensor-core kernels only: CAPTRA_MLP_IMPL=0 (exact-fp32 SA / FP MLPs, a
        # testing knob) keeps the heads on the default tcgen05 variant
        himpl = _mlp.DEFAULT_IMPL if _mlp.DEFAULT_IMPL in (1, 2) else 2
        packs = self._cache.get(self, lambda: [PackedMLP([c.weight.detach().reshape(c.out_channels, c.in_channels)],
                                                         [c.bias.detach()], relu_last=False, impl=himpl) for c in convs])
        x = feat_pm.reshape(B * N, C)
        if N % 128 != 0:      # a 128-row tile would straddle clouds: statistics pass + in-place GroupNorm/ReLU + plain rows
            y = packs[0].rows(x)
            for i in range(1, len(convs)):
                scale, shift = group_norm_affine(y, B, N, gns[i - 1])
                _lib.call("group_norm_relu_rows[R=%d,C=%d]" % (y.shape[0], y.shape[1]), _lib.load().captra_group_norm_relu_rows,
                          y.shape[0], y.shape[1], N, y.data_ptr(), y.stride(0), scale.data_ptr(), shift.data_ptr(),
                          _lib.stream_ptr(y.device), device=y.device)
                y = packs[i].rows(y)
            return y.view(B, N, -1)
        if os.environ.get("CAPTRA_GN_FUSED", "1") == "0":     # unfused statistics pass (A/B knob)
            y = packs[0].rows(x)
            for i in range(1, len(convs)):
                scale, shift = group_norm_affine(y, B, N, gns[i - 1])
                y = packs[i].rows_affine(y, scale, shift, N)
            return y.view(B, N, -1)
        # the statistics of a layer's output are taken in that layer's epilogue (rows_stats) and finalised on
        # 16 MB of block partials, so an activation is written once and read once
        y, stats = packs[0].rows_stats(x, None, None, N)
        for i in range(1, len(convs)):
            scale, shift = group_norm_finalize(stats, B, N, gns[i - 1])
            if i < len(convs) - 1:
                y, stats = packs[i].rows_stats(y, scale, shift, N)
            else:
                y = packs[i].rows_affine(y, scale, shift, N)
        return y.view(B, N, -1)


class RotationRegressor(nn.Module):
    """blocks.py:168-193."""

    def __init__(self, in_dim, num_parts, symmetric=False):
        super().__init__()
        self.sym = symmetric
        rot_dim = 3 if symmetric else 6
        self.rtvec_head = nn.ModuleList([MLPConv1d(in_dim, [512, 512, 256, rot_dim]) for _ in range(num_parts)])
        self.num_parts = num_parts

    def post(self, rtvec):
        """[..., D, N] raw head output -> per-point unit vector (sym) or 3x3 (9 rows)."""
        raw = rtvec.transpose(-1, -2)
        shape = raw.shape
        if self.sym:
            return normalize_vector(raw.reshape(-1, 3)).reshape(shape).transpose(-1, -2)
        rot = compute_rotation_matrix_from_ortho6d(raw.reshape(-1, 6))
        return rot.reshape(shape[:-1] + (-1,)).transpose(-1, -2)

    def forward(self, feat):  # [B, in_dim, N] -> [B, P, R, N]
        return self.post(torch.stack([head(feat) for head in self.rtvec_head], dim=1))


class CoordNet(nn.Module):
    """networks.py:19-110 (tracking-time forward: no 'gt_part' in the input, model.py:349-362)."""

    def __init__(self, cfg):
        super().__init__()
        self.backbone = PointNet2Msg(cfg, cfg['network']['backbone_out_dim'], net_type='camera', use_xyz_feat=True)
        in_dim = cfg['network']['backbone_out_dim']
        self.num_parts = cfg['num_parts']
        self.sym = cfg['obj_sym']
        seg_dim = self.num_parts + cfg['obj']['extra_dims']
        self.seg_head = get_point_mlp(in_dim, seg_dim, [], acti='none')
        self.nocs_head = get_point_mlp(in_dim, 3 * self.num_parts, cfg['network']['nocs_head_dims'], acti='sigmoid')

    def _packed_heads(self):
        def build():
            seg = PackedMLP([self.seg_head[0].weight.detach().reshape(self.seg_head[0].out_channels, -1)],
                            [self.seg_head[0].bias.detach()], relu_last=False)
            convs = [m for m in self.nocs_head if isinstance(m, nn.Conv1d)]
            bns = [m for m in self.nocs_head if isinstance(m, nn.BatchNorm1d)]
            wb = [fold_conv_bn(c, bns[i] if i < len(bns) else None) for i, c in enumerate(convs)]
            nocs = PackedMLP([w for w, _ in wb], [b for _, b in wb], relu_last=False)
            return seg, nocs
        if not hasattr(self, "_cache"):
            self._cache = _FusedCache()
        return self._cache.get(self, build)

    def _pose_branch(self, input, seg, nocs, test):
        """networks.py:54-108: the training-time scale / translation fit from the predicted NOCS with the ground-truth
        rotation -- differentiable through scale_pts_mask / translate_pts_mask, 2-D rotation detached (as in the
        reference); torch ops on the device, the 2x2 SVD on the device kernel instead of the CPU."""
        P = self.num_parts
        pred_labels = torch.argmax(seg, dim=-2)
        labels = pred_labels if test else input['labels']
        rotation = input['gt_part']['rotation']
        final_pose = {'rotation': rotation}
        pred_npcs = nocs.reshape(len(nocs), P, 3, -1)
        cam_points = (input['points'] + input['points_mean']).unsqueeze(1).repeat(1, P, 1, 1)
        eye = torch.cat([torch.eye(P), torch.zeros(2, P)], dim=0).to(pred_npcs.device)
        mask = eye[labels, ].transpose(-1, -2)                       # [B, N, P] -> [B, P, N]
        valid_mask = (mask.sum(dim=-1) > 0).float()
        init_part = input['init_part']
        if self.sym:
            canon_cam = torch.matmul(rotation.transpose(-1, -2), cam_points)
            src_2d = pred_npcs[..., [0, 2], :].transpose(-1, -2)
            tgt_2d = canon_cam[..., [0, 2], :].transpose(-1, -2)
            rot_2d, _ = transform_pts_2d_mask(src_2d, tgt_2d, mask.unsqueeze(-1))
            rotated_npcs = torch.matmul(rotation, torch.matmul(rot_around_yaxis_to_3d(rot_2d), pred_npcs))
        else:
            rotated_npcs = torch.matmul(rotation, pred_npcs)
        scale_mask = mask.unsqueeze(-2)                              # [B, P, 1, N]

        def center(source, m):
            c = torch.sum(source * m, dim=-1, keepdim=True) / torch.clamp(torch.sum(m, dim=-1, keepdim=True), min=1.0)
            return (source - c.detach()) * m

        scale = scale_pts_mask(center(rotated_npcs, scale_mask), center(cam_points, scale_mask), scale_mask)
        scale = valid_mask * scale + (1.0 - valid_mask) * init_part['scale']
        bad = torch.logical_or(torch.isnan(scale), torch.isinf(scale)).float()
        final_pose['scale'] = (1.0 - bad) * scale + bad * init_part['scale']
        used = final_pose['scale'] if test else input['gt_part']['scale']
        scaled_npcs = used.unsqueeze(-1).unsqueeze(-1) * rotated_npcs
        trans = translate_pts_mask(scaled_npcs, cam_points, mask.unsqueeze(-1))
        v = valid_mask.unsqueeze(-1).unsqueeze(-1)
        trans = v * trans + (1.0 - v) * init_part['translation']
        s = trans.sum((-1, -2))
        bad = torch.logical_or(torch.isnan(s), torch.isinf(s)).float().unsqueeze(-1).unsqueeze(-1)
        final_pose['translation'] = (1.0 - bad) * trans + bad * init_part['translation']
        return final_pose

    def forward(self, input, test=False):
        if 'gt_part' not in input and not _needs_autograd(self, input['points']):
            pose = input['canon_pose']
            geom = input.get('geom')
            xyz_pm, cam, dup = frame_ops.canonicalize(input['points'], input['points_mean'], pose['rotation'],
                                                      pose['translation'], pose['scale'], want_cm=True, want_dup=True)
            if geom is not None:
                geom['xyz_pm'] = xyz_pm      # a rigid object's RotationNet sees the very same canonical cloud
            feat = self.backbone.forward_pm(None, geom=geom, xyz_pm=xyz_pm, skip_pm=dup)   # [B,N,C] point-major
            B, N, C = feat.shape
            seg_mlp, nocs_mlp = self._packed_heads()
            x = feat.reshape(B * N, C)
            labels, nocs, seg = frame_ops.coord_head_post(seg_mlp.rows(x), nocs_mlp.rows(x), B, N)
            # 'labels' (= torch.max(seg, dim=-2)[1], model.py:458) is an extra key for the tracker
            return {'seg': seg, 'nocs': nocs, 'points': cam, 'labels': labels}
        cam = canonicalize(input['points'], input['points_mean'], input['canon_pose'])
        feat = self.backbone(cam)
        seg = F.softmax(self.seg_head(feat), dim=1)
        nocs = self.nocs_head(feat) - 0.5
        pred = {'seg': seg, 'nocs': nocs, 'points': cam}
        if 'gt_part' in input:                 # compute s, t from the predicted NOCS (training / evaluation of CoordNet alone)
            pred['part'] = self._pose_branch(input, seg, nocs, test)
        return pred


class RotationRegressionBackbone(nn.Module):
    """networks.py:113-141."""

    def __init__(self, cfg):
        super().__init__()
        self.num_parts = cfg['num_parts']
        self.encoder = PointNet2Msg(cfg, cfg['network']['backbone_out_dim'], use_xyz_feat=False)
        self.sym = cfg['obj_sym']
        self.pose_pred = RotationRegressor(cfg['network']['backbone_out_dim'], self.num_parts, symmetric=self.sym)

    def forward_rotation(self, xyz_pm, labels, rot_prev, batch_size, geom=None, want_rtvec=False):
        """Fused inference path: xyz_pm [B*P,N,3] (copy p canonicalised by part p), labels [B,N], rot_prev [B,P,3,3]
        -> rotation [B,P,3,3] = rot_prev . dR: encoder, head p on copy p (networks.py:200-203 keeps that diagonal only),
        then ONE launch for per-point 6-D / 3-D -> matrix, masked mean, default, Gram-Schmidt and composition."""
        raws = self.forward_heads(xyz_pm, batch_size, geom=geom)
        return frame_ops.rot_head_post(raws, labels, rot_prev, self.sym, want_rtvec=want_rtvec)

    def forward_heads(self, xyz_pm, batch_size, geom=None):
        """Encoder + head p on copy p -> P raw per-point outputs [B,N,D] (point-major): everything of the rotation
        network that does NOT depend on the CoordNet's labels, so it can run on a second stream beside the CoordNet."""
        P = self.num_parts
        feat_pm = self.encoder.forward_pm(None, geom=geom, xyz_pm=xyz_pm)      # [B*P, N, C]
        feat_pm = feat_pm.reshape(batch_size, P, feat_pm.shape[1], feat_pm.shape[2])
        return [self.pose_pred.rtvec_head[p].forward_pm(feat_pm[:, p] if P == 1 else feat_pm[:, p].contiguous()) for p in range(P)]

    def forward_diag(self, cam, labels, batch_size):
        """Autograd / training-mode path in torch ops, as the reference composes it: cam [B*P,3,N] (copy p
        canonicalised by part p), labels [B,N] -> rtvec [B,P,D]: head p on copy p, masked mean over part p's points
        (networks.py:127-139 restricted to the diagonal that networks.py:200-203 keeps)."""
        P = self.num_parts
        feat = self.encoder(cam)                               # [B*P, C, N]
        feat = feat.reshape(batch_size, P, feat.shape[1], feat.shape[2])
        out = []
        for p in range(P):
            raw = self.pose_pred.post(self.pose_pred.rtvec_head[p](feat[:, p]))    # [B, D, N]
            mask = (labels == p).float().unsqueeze(1)                               # [B, 1, N]
            cnt = mask.sum(-1)
            mean = (raw * mask).sum(-1) / torch.clamp_min(cnt, 1.0)
            default = torch.tensor((0., 1., 0.) if self.sym else (1., 0., 0., 0., 1., 0., 0., 0., 1.), device=raw.device)
            valid = (cnt > 0).float()
            out.append(valid * mean + (1.0 - valid) * default.reshape(1, -1))
        return torch.stack(out, dim=1)


class PartCanonNet(nn.Module):
    """networks.py:144-239, network type 'rot_coord_track', test_mode=True (the tracker's call)."""

    def __init__(self, cfg):
        super().__init__()
        self.type = cfg['network']['type']
        assert self.type == 'rot_coord_track', "only the tracking network type is mirrored"
        self.regress_net = RotationRegressionBackbone(cfg)
        self.device = cfg['device']
        self.num_parts = cfg['num_parts']
        self.sym = cfg['obj_sym']
        self.tree = cfg['obj_tree']
        self.root = [i for i in range(self.num_parts) if self.tree[i] == -1][0]
        self.cfg = cfg

    def forward(self, input, test_mode=True):
        assert test_mode, "training-time branches are not mirrored"
        part_pose = input['state']['part']
        P = self.num_parts
        canon_pose = input.get('canon_pose') or {
            key: part_pose[key].reshape((-1,) + part_pose[key].shape[2:]) for key in ('rotation', 'translation', 'scale')}
        cam = input['points']                       # [B,3,N]
        B = len(cam)
        points_mean = input['points_mean']          # [B,3,1]
        labels = input['pred_labels']               # [B,N]
        if not _needs_autograd(self, cam):
            shared = input.get('geom') if P == 1 and 'canon_pose' not in input else None
            if 'rot_raws' in input:
                xyz_pm = None
            elif shared is not None and 'xyz_pm' in shared:
                xyz_pm = shared['xyz_pm']            # rigid object: CoordNet canonicalised by the same pose
            else:
                xyz_pm = frame_ops.canonicalize(cam, points_mean, canon_pose['rotation'], canon_pose['translation'],
                                                canon_pose['scale'], parts=P)[0]
            if 'rot_raws' in input:      # the tracker already ran the label-independent part on its second stream
                rotation = frame_ops.rot_head_post(input['rot_raws'], labels, part_pose['rotation'], self.sym)
            else:
                rotation = self.regress_net.forward_rotation(xyz_pm, labels, part_pose['rotation'], B, geom=shared)
            pred_npcs = input['pred_nocs'].reshape(B, P, 3, -1)
            scale, translation, _ = frame_ops.part_fit_track(labels, pred_npcs.contiguous(), cam, points_mean, rotation, self.sym,
                                                             part_pose['scale'], part_pose['translation'])
            return {'part': {'rotation': rotation, 'scale': scale, 'translation': translation}}
        cam_rep = cam.unsqueeze(1).expand(-1, P, -1, -1).reshape((-1,) + cam.shape[-2:])
        mean_rep = points_mean.unsqueeze(1).expand(-1, P, -1, -1).reshape((-1,) + points_mean.shape[-2:])
        cam_rep = canonicalize(cam_rep, mean_rep, canon_pose)
        rtvec = self.regress_net.forward_diag(cam_rep, labels, B)               # [B,P,D]
        delta_rot = convert_pred_rtvec_to_matrix(rtvec, self.sym)                # [B,P,3,3]
        rotation = torch.matmul(part_pose['rotation'], delta_rot)                # part_dof_utils.py:124-128
        pred_npcs = input['pred_nocs'].reshape(B, P, 3, -1)
        cam_points = (cam + points_mean).unsqueeze(1).expand(-1, P, -1, -1)
        final_pose, valid = part_fit_st_no_ransac(labels, pred_npcs.transpose(-1, -2), cam_points.transpose(-1, -2),
                                                  rotation, {'num_parts': P, 'sym': self.sym})
        v = valid.float()
        final_pose['scale'] = v * final_pose['scale'] + (1.0 - v) * part_pose['scale']
        v = v.unsqueeze(-1).unsqueeze(-1)
        final_pose['translation'] = v * final_pose['translation'] + (1.0 - v) * part_pose['translation']
        return {'part': final_pose}
