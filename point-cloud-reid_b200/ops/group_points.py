"""grouping_operation / QueryAndGroup / GroupAll.
Reference: mmdet3d/ops/group_points/group_points.py:11-129 (QueryAndGroup), 132-166 (GroupAll),
169-220 (GroupingOperation; forward only here)."""
import torch
from torch import nn

from ._common import _NoBackward, check, lib, ptr, require, stream
from .ball_query import ball_query
from .knn import knn


class GroupingOperation(_NoBackward):
    @staticmethod
    def forward(ctx, features, indices):
        require(features, "features")
        require(indices, "indices", torch.int32)
        B, nfeatures, nsample = indices.shape
        _, C, N = features.shape
        with torch.cuda.device(features.device):
            output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
            check(lib().pcreid_group_points(B, C, N, nfeatures, nsample, ptr(features), ptr(indices), ptr(output), stream()),
                  "pcreid_group_points")
        return output


grouping_operation = GroupingOperation.apply


class QueryAndGroup(nn.Module):
    """kNN (max_radius None) or ball query, then group xyz offsets and features channel-first."""

    def __init__(self, max_radius, sample_num, min_radius=0, use_xyz=True, return_grouped_xyz=False, normalize_xyz=False,
                 uniform_sample=False, return_unique_cnt=False, return_grouped_idx=False):
        super().__init__()
        self.max_radius = max_radius
        self.min_radius = min_radius
        self.sample_num = sample_num
        self.use_xyz = use_xyz
        self.return_grouped_xyz = return_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.uniform_sample = uniform_sample
        self.return_unique_cnt = return_unique_cnt
        self.return_grouped_idx = return_grouped_idx
        if self.return_unique_cnt:
            assert self.uniform_sample, "uniform_sample should be True when returning the count of unique samples"
        if self.max_radius is None:
            assert not self.normalize_xyz, "can not normalize grouped xyz when max_radius is None"

    def forward(self, points_xyz, center_xyz, features=None):
        if self.max_radius is None:
            idx = knn(self.sample_num, points_xyz, center_xyz, False).transpose(1, 2).contiguous()
        else:
            idx = ball_query(self.min_radius, self.max_radius, self.sample_num, points_xyz, center_xyz)
        unique_cnt = None
        if self.uniform_sample:
            unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
            for i_batch in range(idx.shape[0]):
                for i_region in range(idx.shape[1]):
                    unique_ind = torch.unique(idx[i_batch, i_region, :])
                    num_unique = unique_ind.shape[0]
                    unique_cnt[i_batch, i_region] = num_unique
                    sample_ind = torch.randint(0, num_unique, (self.sample_num - num_unique,), dtype=torch.long)
                    idx[i_batch, i_region, :] = torch.cat((unique_ind, unique_ind[sample_ind]))
        xyz_trans = points_xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz_diff = grouped_xyz - center_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz_diff /= self.max_radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz_diff, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz_diff
        ret = [new_features]
        if self.return_grouped_xyz:
            ret.append(grouped_xyz)
        if self.return_unique_cnt:
            ret.append(unique_cnt)
        if self.return_grouped_idx:
            ret.append(idx)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz
