"""Mirror of network/models/pointnet_utils.py: the op shims and the three live PointNet++ modules,
with identical constructor signatures, forward signatures, tensor layouts ([B,C,N]) and
state-dict keys (conv_blocks.i.j / bn_blocks.i.j / mlp_convs.i / mlp_bns.i), so reference
checkpoints load unchanged.

Eval-mode forwards run on the fused kernels (one launch per SA scale / FP stage, BN folded);
training-mode forwards (BatchNorm batch statistics, autograd) use the same composition as the
reference on top of this package's CUDA ops -- there is no CPU path anywhere.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused_ops
from .mlp import PackedMLP, fold_conv_bn
from .pointnet_lib import pointnet2_utils as futils

CUDA = True  # the reference's flag (pointnet_utils.py:8); this package is CUDA-only


# ---- op shims (pointnet_utils.py:12-168) ------------------------------------------------------
def knn_point(k, pos2, pos1):
    """pointnet_utils.py:12-32."""
    val, idx = futils.knn(k, pos2, pos1)
    return val, idx.long()


def three_nn(xyz1, xyz2):
    """pointnet_utils.py:35-43."""
    dists, idx = futils.three_nn(xyz1, xyz2)
    return dists, idx.long()


def three_interpolate(points, idx, weight):
    """pointnet_utils.py:46-55.  points [B,C,M], idx [B,N,3] -> [B,C,N]."""
    return futils.three_interpolate(points, idx.int(), weight)


def square_distance(src, dst):
    """pointnet_utils.py:58-79."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def index_points(points, idx):
    """pointnet_utils.py:82-97.  points [B,N,C], idx [B,S,...] -> [B,S,...,C]."""
    B = points.shape[0]
    view = [B] + [1] * (idx.dim() - 1)
    batch = torch.arange(B, dtype=torch.long, device=points.device).view(view).expand_as(idx)
    return points[batch, idx, :]


def gather_operation(feature, idx):
    """pointnet_utils.py:100-103 ([B,C,N],[B,S] -> [B,C,S]) on the gather kernel."""
    return futils.gather_operation(feature, idx.int())


def group_operation(feature, idx):
    """pointnet_utils.py:106-109 ([B,C,N],[B,S,K] -> [B,C,S,K]) on the grouping kernel."""
    return futils.grouping_operation(feature, idx.int())


def farthest_point_sample(xyz, npoint):
    """pointnet_utils.py:112-138 (CUDA branch :124): [B,N,3] -> [B,npoint] int64."""
    return futils.furthest_point_sample(xyz, npoint).long()


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet_utils.py:141-168 (CUDA branch :155): -> [B,S,nsample] int64."""
    return futils.ball_query(radius, nsample, xyz, new_xyz).long()


def sample_and_group_all(xyz, points):
    """pointnet_utils.py:171-188: channels [xyz, points]."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        return new_xyz, torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped_xyz


# ---- fused-path plumbing -------------------------------------------------------------------
def _params_key(module):
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


class _FusedCache:
    """Packs (and re-packs when any parameter/buffer changed) the folded weights of a module."""

    def __init__(self):
        self._key, self._val = None, None

    def get(self, module, builder):
        key = _params_key(module)
        if key != self._key:
            self._val, self._key = builder(), key
        return self._val


def _needs_autograd(module, *tensors):
    if module.training:
        return True
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())


class SharedGeom(dict):
    """Coordinate-only results (FPS picks, ball-query lists, 3-NN indices / weights, the canonicalised cloud) shared by
    two networks that run on DIFFERENT CUDA streams: storing an item records an event on the producing stream, reading it
    from another stream makes that stream wait for the event.  The producer's launches must be enqueued (host side)
    before the consumer asks; on the device the two chains then overlap wherever the data dependencies allow."""

    def __init__(self, aux=None):
        super().__init__()
        self._ev = {}
        self.aux = aux          # optional helper stream: the second SA level's sampling / grouping runs there, beside the first level's MLPs

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        ev = torch.cuda.Event()
        st = torch.cuda.current_stream()
        ev.record(st)
        self._ev[key] = (ev, st)

    def __getitem__(self, key):
        value = super().__getitem__(key)
        rec = self._ev.get(key)
        if rec is not None:
            cur = torch.cuda.current_stream()
            if cur != rec[1]:
                cur.wait_event(rec[0])
        return value

    def setdefault(self, key, default=None):
        if key not in self:
            super().__setitem__(key, SharedGeom(self.aux) if isinstance(default, dict) and not default else default)
        return super().__getitem__(key)


def _pm(t):
    """[B,C,N] -> point-major contiguous [B,N,C]."""
    return t.transpose(1, 2).contiguous()


def split_first_layer(W0, b0, n_feat):
    """Split the BN-folded first layer of an SA scale, y = W0 [features | x - c] + b0 (the reference concatenates
    the grouped features first, then the centred coordinates: pointnet_utils.py:239-241), into the per-point part
    and the per-(centroid, sample) part of captra_sa_mlp_max_pre:
        Wf  [c1, n_feat]  -- applied once per point:  P[j] = Wf f_j
        tab [4, c1]       -- rows wx, wy, wz, bias:   y = P[j] + wx dx + wy dy + wz dz + bias."""
    Wf = W0[:, :n_feat]
    tab = torch.cat([W0[:, n_feat:n_feat + 3].t(), b0[None, :]], 0).contiguous()
    return Wf, tab


class PointNetSetAbstractionMsg(nn.Module):
    """pointnet_utils.py:191-250."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, knn=False):
        super().__init__()
        self.npoint = npoint
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        self.out_channel = 0
        self.scale_channels = []
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel
            for out_channel in mlp:
                convs.append(nn.Conv2d(last, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last = out_channel
            self.out_channel += last
            self.scale_channels.append(last)
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.knn = knn
        self._cache = _FusedCache()
        self._cache_pre = _FusedCache()

    # Input feature widths from which layer 0 is projected per point instead of per (centroid, sample) row
    # (PackedMLP.sa_max_pre): worthwhile once a point's feature row is much wider than its 3 coordinates.
    PRE_MIN_FEAT = 32

    def _packed(self):
        def build():
            out = []
            for convs, bns in zip(self.conv_blocks, self.bn_blocks):
                wb = [fold_conv_bn(c, b) for c, b in zip(convs, bns)]
                out.append(PackedMLP([w for w, _ in wb], [b for _, b in wb], relu_last=True))
            return out
        return self._cache.get(self, build)

    def _packed_pre(self):
        """Projected form of every scale: (projector over all scales, [(tail mlp, coordinate/bias table, column
        offset, width)]) or None when the shapes do not qualify (few input features, one-layer scales, the
        exact-fp32 kernel selected, first layers wider than one 256-column launch together)."""
        def build():
            from . import mlp as _mlp
            if _mlp.DEFAULT_IMPL not in (1, 2) or os.environ.get("CAPTRA_SA_PRE", "1") == "0":   # knob: A/B timing
                return None
            folded = [[fold_conv_bn(c, b) for c, b in zip(convs, bns)] for convs, bns in zip(self.conv_blocks, self.bn_blocks)]
            D = folded[0][0][0].shape[1] - 3
            if D < self.PRE_MIN_FEAT or any(len(wb) < 2 for wb in folded) or sum(wb[0][0].shape[0] for wb in folded) > 256:
                return None
            scales, off = [], 0
            for wb in folded:
                W0, b0 = wb[0]
                tail = PackedMLP([w for w, _ in wb[1:]], [b for _, b in wb[1:]], relu_last=True)
                if tail._layers is not None or not tail._tc_supported():
                    return None
                tab = split_first_layer(W0, b0, D)[1]                                   # [4, c1]: wx, wy, wz, bias
                scales.append((tail, tab, off, W0.shape[0]))
                off += W0.shape[0]
            Wf = torch.cat([split_first_layer(wb[0][0], wb[0][1], D)[0] for wb in folded], 0).contiguous()   # [sum c1, D]
            proj = PackedMLP([Wf], [torch.zeros(Wf.shape[0], device=Wf.device)], relu_last=False)
            return proj, scales
        return self._cache_pre.get(self, build)

    def geometry(self, xyz_pm):
        """The coordinate-only part of the layer (pointnet_utils.py:225-233): xyz [B,N,3] -> (new_xyz [B,S,3], [idx [B,S,K_r]])."""
        if xyz_pm.shape[1] <= 8192 and os.environ.get("CAPTRA_FPS_BQ_PIPE", "1") != "0":   # knob: A/B timing
            # sampling and grouping indices as a pipeline: the ball query consumes centroids while FPS still picks
            return fused_ops.fps_ball_query(xyz_pm, self.npoint, self.radius_list, self.nsample_list)
        _, new_xyz = fused_ops.fps_gather(xyz_pm, self.npoint)
        return new_xyz, fused_ops.ball_query_multi(self.radius_list, self.nsample_list, xyz_pm, new_xyz)

    def forward_pm(self, xyz_pm, feats_pm, geom=None):
        """Fused inference path on point-major tensors: xyz [B,N,3], feats [B,N,D] or None ->
        (new_xyz [B,S,3], new_feats [B,S,sum(cout)]).  `geom` (a dict) carries the sampling and
        grouping indices: filled on the first call, reused when another network runs on the very
        same coordinates (CoordNet / RotationNet of a rigid object)."""
        assert not self.knn, "knn grouping is dead in the reference (knn=False everywhere)"
        if feats_pm is not None and feats_pm.shape[-1] == 0:
            feats_pm = None
        if geom is not None and "new_xyz" in geom:
            new_xyz, idxs = geom["new_xyz"], geom["idxs"]
        else:
            new_xyz, idxs = self.geometry(xyz_pm)
            if geom is not None:
                geom["new_xyz"], geom["idxs"] = new_xyz, idxs
        out = torch.empty(xyz_pm.shape[0], self.npoint, self.out_channel, dtype=torch.float32, device=xyz_pm.device)
        off = 0
        pre = self._packed_pre() if (feats_pm is not None and all(i.shape[2] in PackedMLP.TC_GROUPS for i in idxs)) else None
        if pre is not None:
            # layer 0 of all scales projected once per point (one dense launch), then gathered: the feature part
            # of the first conv depends on the point only, and a point sits in nsample * S / N balls
            proj, scales = pre
            B, N, D = feats_pm.shape
            P = proj.rows(feats_pm.reshape(B * N, D))
            for (tail, tab, poff, c1), idx in zip(scales, idxs):
                tail.sa_max_pre(xyz_pm, new_xyz, P[:, poff:poff + c1], tab, idx, out, col_off=off)
                off += tail.cout
            return new_xyz, out
        for mlp, idx in zip(self._packed(), idxs):
            mlp.sa_max(xyz_pm, new_xyz, feats_pm, idx, out, col_off=off)
            off += mlp.cout
        return new_xyz, out

    def forward(self, xyz, points):
        """xyz [B,C,N], points [B,D,N] -> (new_xyz [B,C,S], new_points [B,D',S])."""
        if not _needs_autograd(self, xyz, points):
            new_xyz, out = self.forward_pm(_pm(xyz), _pm(points) if points is not None else None)
            return new_xyz.transpose(1, 2), out.transpose(1, 2)
        B, C, N = xyz.shape
        S = self.npoint
        fps_idx = farthest_point_sample(xyz.permute(0, 2, 1), S)
        new_xyz = gather_operation(xyz, fps_idx)
        new_points_list = []
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            if self.knn:
                _, group_idx = knn_point(K, new_xyz.transpose(-1, -2), xyz.transpose(-1, -2))
            else:
                group_idx = query_ball_point(radius, K, xyz.transpose(-1, -2), new_xyz.transpose(-1, -2))
            grouped_xyz = group_operation(xyz, group_idx) - new_xyz.view(B, C, S, 1)
            if points is not None:
                grouped_points = torch.cat([group_operation(points, group_idx), grouped_xyz], dim=1)
            else:
                grouped_points = grouped_xyz
            for conv, bn in zip(self.conv_blocks[i], self.bn_blocks[i]):
                grouped_points = F.relu(bn(conv(grouped_points)))
            new_points_list.append(torch.max(grouped_points, -1)[0])
        return new_xyz, torch.cat(new_points_list, dim=1)


class PointNetFeaturePropagation(nn.Module):
    """pointnet_utils.py:253-299."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel
        self.out_channel = last
        self._cache = _FusedCache()

    def folded(self):
        return [fold_conv_bn(c, b) for c, b in zip(self.mlp_convs, self.mlp_bns)]

    def _packed(self):
        def build():
            wb = self.folded()
            return PackedMLP([w for w, _ in wb], [b for _, b in wb], relu_last=True)
        return self._cache.get(self, build)

    def forward_pm(self, xyz1_pm, xyz2_pm, points1_pm, points2_pm, mlp=None, geom=None):
        """Fused inference path: xyz1 [B,N,3], xyz2 [B,S,3], points1 [B,N,D1] or None,
        points2 [B,S,D2] -> [B,N,cout].  `mlp` lets the caller append layers (backbones.py:68);
        `geom` caches the 3-NN indices/weights for a second network on the same coordinates."""
        B, N, _ = xyz1_pm.shape
        S = points2_pm.shape[1]
        mlp = mlp or self._packed()
        segA = points1_pm.reshape(B * N, -1) if points1_pm is not None and points1_pm.shape[-1] > 0 else None
        if S == 1:  # pointnet_utils.py:281-282: repeat the single coarse point
            out = mlp.rows(segA, points2_pm.reshape(B, -1), bcast_rows=N)
        else:
            if geom is not None and "nn" in geom:
                interp = fused_ops.three_nn_interpolate_pm(xyz1_pm, xyz2_pm, points2_pm, nn=geom["nn"])
            elif geom is not None:
                interp, geom["nn"] = fused_ops.three_nn_interpolate_pm(xyz1_pm, xyz2_pm, points2_pm, return_nn=True)
            else:
                interp = fused_ops.three_nn_interpolate_pm(xyz1_pm, xyz2_pm, points2_pm)
            out = mlp.rows(segA, interp.reshape(B * N, -1))
        return out.view(B, N, -1)

    def forward(self, xyz1, xyz2, points1, points2):
        """xyz1 [B,C,N], xyz2 [B,C,S], points1 [B,D,N], points2 [B,D,S] -> [B,D',N]."""
        if not _needs_autograd(self, xyz1, xyz2, points1, points2):
            out = self.forward_pm(_pm(xyz1), _pm(xyz2), _pm(points1) if points1 is not None else None, _pm(points2))
            return out.transpose(1, 2)
        xyz1 = xyz1.permute(0, 2, 1)
        xyz2 = xyz2.permute(0, 2, 1)
        B, N, C = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated_points = points2.repeat(1, 1, N)
        else:
            dist, idx = three_nn(xyz1, xyz2)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated_points = three_interpolate(points2, idx, weight)
        new_points = torch.cat([points1, interpolated_points], dim=-2) if points1 is not None else interpolated_points
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_points = F.relu(bn(conv(new_points)))
        return new_points


class PointNetSetAbstraction(nn.Module):
    """pointnet_utils.py:302-343 (group_all only, as in the reference)."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, knn=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel
        self.out_channel = last
        self.group_all = group_all
        self.knn = knn
        self._cache = _FusedCache()

    def _packed(self):
        def build():
            wb = [fold_conv_bn(c, b) for c, b in zip(self.mlp_convs, self.mlp_bns)]
            return PackedMLP([w for w, _ in wb], [b for _, b in wb], relu_last=True)
        return self._cache.get(self, build)

    def forward_pm(self, xyz_pm, feats_pm):
        """xyz [B,N,3], feats [B,N,D] or None -> [B,cout]: channels [xyz, feats] (:185), max over N."""
        assert self.group_all, "Not Implemented"  # pointnet_utils.py:335
        B, N, _ = xyz_pm.shape
        segB = feats_pm.reshape(B * N, -1) if feats_pm is not None and feats_pm.shape[-1] > 0 else None
        return self._packed().rows(xyz_pm.reshape(B * N, 3), segB, group=N)

    def forward(self, xyz, points):
        """xyz [B,C,N], points [B,D,N] -> (new_xyz [B,C,1] zeros, new_points [B,D',1])."""
        if not _needs_autograd(self, xyz, points):
            out = self.forward_pm(_pm(xyz), _pm(points) if points is not None else None)
            return torch.zeros(xyz.shape[0], xyz.shape[1], 1, device=xyz.device), out.unsqueeze(-1)
        assert self.group_all, "Not Implemented"
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        new_xyz, new_points = sample_and_group_all(xyz, points)
        new_points = new_points.permute(0, 3, 2, 1)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_points = F.relu(bn(conv(new_points)))
        return new_xyz.permute(0, 2, 1), torch.max(new_points, 2)[0]
