"""Host-side mirror of network/models/pointnet_lib/pointnet2_utils.py.

Same public names, argument order, output shapes/dtypes and autograd behaviour as the reference
file (line numbers cited per class), so `from pointnet_lib import pointnet2_utils as futils`
(network/models/pointnet_utils.py:10) keeps working when this package is installed as the
drop-in.  Outputs are allocated here and only written by the kernels, exactly like the
reference (its legacy torch.cuda.FloatTensor(...) allocators become torch.empty on the input's
device).  All compute goes through captra_b200.pointnet2_cuda -> libcaptra_ops.so.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import pointnet2_cuda as pointnet2


def _new(ref, shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=ref.device)


class FurthestPointSampling(Function):
    """pointnet2_utils.py:10-34.  xyz (B,N,3) -> (B,npoint) int32, starts from index 0."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        xyz = xyz.contiguous()
        B, N, _ = xyz.size()
        output = _new(xyz, (B, npoint), torch.int32)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:40-73.  features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, npoint))
        pointnet2.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class KNN(Function):
    """pointnet2_utils.py:78-106.  Returns (sqrt(dist2), idx) with k <= 200."""

    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unknown = unknown.contiguous()
        known = known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, k))
        idx = _new(unknown, (B, N, k), torch.int32)
        pointnet2.knn_wrapper(B, N, m, k, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    """pointnet2_utils.py:110-138.  unknown (B,N,3), known (B,M,3) -> (sqrt dist (B,N,3), idx int32)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unknown = unknown.contiguous()
        known = known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, 3))
        idx = _new(unknown, (B, N, 3), torch.int32)
        pointnet2.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """pointnet2_utils.py:144-189.  features (B,C,M), idx/weight (B,n,3) -> (B,C,n); grad wrt features."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous()
        weight = weight.contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = _new(features, (B, c, n))
        pointnet2.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:195-235.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous().int()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, nfeatures, nsample))
        pointnet2.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:241-268.  NOTE the Python order (radius, nsample, xyz, new_xyz) vs the
    C order (new_xyz, xyz); empty balls keep the zeros written here (:261)."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        new_xyz = new_xyz.contiguous()
        xyz = xyz.contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        pointnet2.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:274-307: ball query + grouping; channels are [features, rel_xyz] (:300)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_features, grouped_xyz], dim=1)
        return grouped_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:310-332: channels are [xyz, features] (:327)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features


class KNNAndGroup(nn.Module):
    """pointnet2_utils.py:335-385.  The reference calls knn() with the wrong arity (:360, dead
    code); here idx defaults to the k=nsample nearest neighbours of new_xyz in xyz."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor = None, idx: torch.Tensor = None,
                features: torch.Tensor = None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            _, idx = knn(self.nsample, new_xyz, xyz)
        idx = idx.detach()
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features
