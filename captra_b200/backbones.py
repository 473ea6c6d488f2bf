"""Mirror of network/models/backbones.py: PointNet2Msg (sa1 -> sa2 -> sa3 -> fp3 -> fp2 -> fp1 ->
conv1+bn1+relu), same constructor, same state-dict keys, same [B,3,N] -> [B,out_dim,N] contract.

Eval-mode forward = 13 kernel launches on point-major tensors: 2x (FPS+gather, multi-radius ball
query), 5 fused SA scales, group-all SA, fp3, 2x (3-NN+interpolate, FP MLP) with conv1/bn1 folded
into the fp1 chain as a third layer.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mlp import PackedMLP, fold_conv_bn
from .pointnet_utils import (PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg,
                             _FusedCache, _needs_autograd)


class PointNet2Msg(nn.Module):
    """backbones.py:15-69.  Output: out_dim channels per point."""

    def __init__(self, cfg, out_dim, net_type='camera', use_xyz_feat=False):
        super().__init__()
        net_cfg = cfg['pointnet'][net_type]
        self.out_dim = out_dim
        self.in_dim = 3 if use_xyz_feat else 0
        self.use_xyz_feat = use_xyz_feat
        self.sa1 = PointNetSetAbstractionMsg(npoint=net_cfg['sa1']['npoint'],
                                             radius_list=net_cfg['sa1']['radius_list'],
                                             nsample_list=net_cfg['sa1']['nsample_list'],
                                             in_channel=self.in_dim + 3,
                                             mlp_list=net_cfg['sa1']['mlp_list'])
        self.sa2 = PointNetSetAbstractionMsg(npoint=net_cfg['sa2']['npoint'],
                                             radius_list=net_cfg['sa2']['radius_list'],
                                             nsample_list=net_cfg['sa2']['nsample_list'],
                                             in_channel=self.sa1.out_channel + 3,
                                             mlp_list=net_cfg['sa2']['mlp_list'])
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None,
                                          in_channel=self.sa2.out_channel + 3,
                                          mlp=net_cfg['sa3']['mlp'], group_all=True)
        self.fp3 = PointNetFeaturePropagation(in_channel=self.sa2.out_channel + self.sa3.out_channel,
                                              mlp=net_cfg['fp3']['mlp'])
        self.fp2 = PointNetFeaturePropagation(in_channel=self.sa1.out_channel + self.fp3.out_channel,
                                              mlp=net_cfg['fp2']['mlp'])
        self.fp1 = PointNetFeaturePropagation(in_channel=self.in_dim + 3 + self.fp2.out_channel,
                                              mlp=net_cfg['fp1']['mlp'])
        self.conv1 = nn.Conv1d(self.fp1.out_channel, self.out_dim, 1)
        self.bn1 = nn.BatchNorm1d(self.out_dim)
        self.device = cfg['device']
        self._cache = _FusedCache()

    def _fp1_head(self):
        """fp1's two layers + conv1/bn1/relu (backbones.py:68) as one 3-layer chain."""
        def build():
            wb = self.fp1.folded() + [fold_conv_bn(self.conv1, self.bn1)]
            return PackedMLP([w for w, _ in wb], [b for _, b in wb], relu_last=True)
        return self._cache.get(self, build)

    def forward_pm(self, input, geom=None, xyz_pm=None, skip_pm=None):
        """input [B,3,N] -> feat [B,N,out_dim] (point-major), fused inference path.  The caller may hand over the
        cloud already point-major (`xyz_pm` [B,N,3], and `skip_pm` [B,N,6] = [xyz, xyz] when use_xyz_feat; both are
        by-products of captra_canonicalize), in which case `input` is not read.  `geom`: an
        (initially empty) dict that receives every coordinate-only result (FPS picks, ball-query
        index lists, 3-NN indices and weights); passing the same dict to a second backbone that is
        evaluated on the *same* input coordinates reuses them (bit-identical, they depend on xyz only)."""
        g = (lambda k: geom.setdefault(k, {})) if geom is not None else (lambda k: None)
        l0_xyz = xyz_pm if xyz_pm is not None else input.transpose(1, 2).contiguous()   # [B,N,3]
        l0_feats = l0_xyz if self.use_xyz_feat else None                   # backbones.py:57-60
        aux = getattr(geom, "aux", None)
        if aux is not None and os.environ.get("CAPTRA_GEOM_AUX", "1") != "0":
            # The second level's sampling + grouping indices depend on the first level's CENTROIDS only, not on its
            # features: run them on a helper stream, beside the first level's MLPs (FPS keeps one small CTA per cloud
            # busy for 128 dependent rounds).  SharedGeom's events order producer and consumers across the streams.
            g1, g2 = g("sa1"), g("sa2")
            if "new_xyz" not in g1:
                g1["new_xyz"], g1["idxs"] = self.sa1.geometry(l0_xyz)
            if "new_xyz" not in g2:
                with torch.cuda.stream(aux):
                    g2["new_xyz"], g2["idxs"] = self.sa2.geometry(g1["new_xyz"])
        l1_xyz, l1_feats = self.sa1.forward_pm(l0_xyz, l0_feats, geom=g("sa1"))
        l2_xyz, l2_feats = self.sa2.forward_pm(l1_xyz, l1_feats, geom=g("sa2"))
        l3_feats = self.sa3.forward_pm(l2_xyz, l2_feats)                   # [B,1024]
        B = l0_xyz.shape[0]
        # fp3 interpolates from the single group-all point: its coordinates (zeros, pointnet_utils.py:176) are never read
        l2_feats = self.fp3.forward_pm(l2_xyz, None, l2_feats, l3_feats.view(B, 1, -1))
        l1_feats = self.fp2.forward_pm(l1_xyz, l2_xyz, l1_feats, l2_feats, geom=g("fp2"))
        if self.use_xyz_feat:                                                          # backbones.py:67
            skip = skip_pm if skip_pm is not None else torch.cat([l0_xyz, l0_xyz], dim=-1)
        else:
            skip = l0_xyz
        return self.fp1.forward_pm(l0_xyz, l1_xyz, skip, l1_feats, mlp=self._fp1_head(), geom=g("fp1"))

    def forward(self, input):  # [B,3,N]
        if not _needs_autograd(self, input):
            return self.forward_pm(input).transpose(1, 2)
        l0_xyz = input
        l0_points = input if self.use_xyz_feat else input[:, 3:]
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(l0_xyz, l1_xyz, torch.cat([l0_xyz, l0_points], dim=1), l1_points)
        return F.relu(self.bn1(self.conv1(l0_points)))
