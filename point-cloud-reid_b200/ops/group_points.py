"""grouping_operation / QueryAndGroup / GroupAll.
Reference: mmdet3d/ops/group_points/group_points.py:11-129 (QueryAndGroup), 132-166 (GroupAll),
169-220 (GroupingOperation; forward only here)."""
import torch
from torch import nn

from ._common import OPS, _NoBackward, require
from .ball_query import ball_query
from .knn import knn


class GroupingOperation(_NoBackward):
    @staticmethod
    def forward(ctx, features, indices):
        require(features, "features")
        require(indices, "indices", torch.int32)
        B, nfeatures, nsample = indices.shape
        _, C, N = features.shape
        output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
        OPS.group_points(B, C, N, nfeatures, nsample, features, indices, output)
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
        """points_xyz (B, N, 3), center_xyz (B, S, 3), features (B, C, N) or None -> (B, 3 + C, S, k) [+ grouped xyz, unique
        counts, indices as configured].  One index query + ONE fused kernel (pcreid_query_group) that writes the centred
        (optionally radius-normalised) neighbour offsets and the gathered features straight into the concatenated result."""
        require(points_xyz, "points_xyz")
        require(center_xyz, "center_xyz")
        if features is not None:
            require(features, "features")
        elif not self.use_xyz:
            raise AssertionError("Cannot have not features and not use xyz as a feature!")
        if self.max_radius is None:
            idx = knn(self.sample_num, points_xyz, center_xyz, False).transpose(1, 2).contiguous()
        else:
            idx = ball_query(self.min_radius, self.max_radius, self.sample_num, points_xyz, center_xyz)
        counts = None
        if self.uniform_sample:
            idx, counts = _spread_unique(idx)
        B, S, k = idx.shape
        n_feat = 0 if features is None else features.shape[1]
        new_features = torch.empty((B, (3 if self.use_xyz else 0) + n_feat, S, k), dtype=torch.float32, device=points_xyz.device)
        grouped_xyz = torch.empty((B, 3, S, k), dtype=torch.float32, device=points_xyz.device) if self.return_grouped_xyz else None
        OPS.query_group(B, n_feat, points_xyz.shape[1], S, k, points_xyz, center_xyz, features, idx, int(self.use_xyz),
                        float(self.max_radius) if self.normalize_xyz else 0.0, new_features, grouped_xyz)
        extras = [t for t, wanted in ((grouped_xyz, self.return_grouped_xyz), (counts, self.return_unique_cnt),
                                      (idx, self.return_grouped_idx)) if wanted]
        return (new_features, *extras) if extras else new_features


def _spread_unique(idx):
    """uniform_sample of the reference (group_points.py:78-91): every region keeps its distinct indices (ascending, as
    torch.unique returns them) and fills the remaining slots with random picks among them.  The distinct sets are found for
    all regions at once on the device (sort, first-occurrence mask, stable compaction); only the random fill is drawn region
    by region on the host generator, with the same call the reference makes, so a seeded run picks the same samples.
    -> (idx (B, S, k) int32, distinct counts (B, S) float32 on the CPU like the reference's)."""
    B, S, k = idx.shape
    srt = idx.sort(dim=2).values
    first = torch.ones_like(srt, dtype=torch.bool)
    first[..., 1:] = srt[..., 1:] != srt[..., :-1]
    distinct = srt.gather(2, (~first).to(torch.uint8).sort(dim=2, stable=True).indices)    # distinct values first, ascending
    n_distinct = first.sum(2).cpu()
    pick = torch.arange(k).repeat(B, S, 1)
    for b, s in ((b, s) for b in range(B) for s in range(S)):
        n = int(n_distinct[b, s])
        if n < k:
            pick[b, s, n:] = torch.randint(0, n, (k - n,), dtype=torch.long)
    return distinct.gather(2, pick.to(idx.device)).to(torch.int32).contiguous(), n_distinct.to(torch.float32)


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
