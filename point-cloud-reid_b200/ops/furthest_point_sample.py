"""furthest_point_sample / furthest_point_sample_with_dist / Points_Sampler.
Reference: mmdet3d/ops/furthest_point_sample/furthest_point_sample.py:7-78, points_sampler.py:11-157."""
import torch
from torch import nn

from ._common import OPS, require


def _fps(points_xyz, num_points, with_dist):
    require(points_xyz, "points_xyz")
    B, N = points_xyz.shape[:2]
    if with_dist:
        assert points_xyz.dim() == 3 and points_xyz.shape[2] == N, "points_dist must be (B, N, N)"
    else:
        assert points_xyz.dim() == 3 and points_xyz.shape[2] == 3, "points_xyz must be (B, N, 3)"
    output = torch.empty((B, num_points), dtype=torch.int32, device=points_xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=points_xyz.device)
    (OPS.fps_with_dist if with_dist else OPS.fps)(B, N, num_points, points_xyz, temp, output)
    return output


class FurthestPointSampling:
    """(B, N, 3) xyz -> (B, num_points) int32 indices; first index is always 0."""

    @staticmethod
    def apply(points_xyz, num_points):
        return _fps(points_xyz, int(num_points), False)

    forward = apply


class FurthestPointSamplingWithDist:
    """(B, N, N) pairwise distances -> (B, num_points) int32 indices."""

    @staticmethod
    def apply(points_dist, num_points):
        return _fps(points_dist, int(num_points), True)

    forward = apply


furthest_point_sample = FurthestPointSampling.apply
furthest_point_sample_with_dist = FurthestPointSamplingWithDist.apply


def calc_square_dist(point_feat_a, point_feat_b, norm=True):
    """furthest_point_sample/utils.py:4-31: pairwise squared distance of (B, N, C) / (B, M, C) features -> (B, N, M), one
    kernel (pcreid_pairwise_sqdist) instead of the reference's sum / repeat / matmul / sqrt chain of torch ops."""
    a = point_feat_a.contiguous().float()
    b = point_feat_b.contiguous().float()
    require(a, "point_feat_a")
    require(b, "point_feat_b")
    assert a.dim() == 3 and b.dim() == 3 and a.shape[0] == b.shape[0] and a.shape[2] == b.shape[2]
    B, N, C = a.shape
    M = b.shape[1]
    dist = torch.empty((B, N, M), dtype=torch.float32, device=a.device)
    OPS.pairwise_sqdist(B, N, M, C, a, b, dist, int(bool(norm)))
    return dist


class DFPS_Sampler(nn.Module):
    def forward(self, points, features, npoint):
        return furthest_point_sample(points.contiguous(), npoint)


class FFPS_Sampler(nn.Module):
    def forward(self, points, features, npoint):
        assert features is not None, "feature input to FFPS_Sampler should not be None"
        features_for_fps = torch.cat([points, features.transpose(1, 2)], dim=2)
        features_dist = calc_square_dist(features_for_fps, features_for_fps, norm=False)
        return furthest_point_sample_with_dist(features_dist.contiguous(), npoint)


class FS_Sampler(nn.Module):
    def forward(self, points, features, npoint):
        assert features is not None, "feature input to FS_Sampler should not be None"
        fidx_ffps = FFPS_Sampler()(points, features, npoint)
        fidx_dfps = DFPS_Sampler()(points, features, npoint)
        return torch.cat([fidx_ffps, fidx_dfps], dim=1)


def get_sampler_type(sampler_type):
    try:
        return {"D-FPS": DFPS_Sampler, "F-FPS": FFPS_Sampler, "FS": FS_Sampler}[sampler_type]
    except KeyError:
        raise ValueError(f'Only "sampler_type" of "D-FPS", "F-FPS", or "FS" are supported, got {sampler_type}')


class Points_Sampler(nn.Module):
    """points_sampler.py:34-104: concatenates the indices of several samplers over point ranges."""

    def __init__(self, num_point, fps_mod_list=["D-FPS"], fps_sample_range_list=[-1]):
        super().__init__()
        assert len(num_point) == len(fps_mod_list) == len(fps_sample_range_list)
        self.num_point = num_point
        self.fps_sample_range_list = fps_sample_range_list
        self.samplers = nn.ModuleList([get_sampler_type(m)() for m in fps_mod_list])

    def forward(self, points_xyz, features):
        indices = []
        last = 0
        for fps_sample_range, sampler, npoint in zip(self.fps_sample_range_list, self.samplers, self.num_point):
            assert fps_sample_range < points_xyz.shape[1]
            if fps_sample_range == -1:
                sample_points_xyz = points_xyz[:, last:]
                sample_features = features[:, :, last:] if features is not None else None
            else:
                sample_points_xyz = points_xyz[:, last:fps_sample_range]
                sample_features = features[:, :, last:fps_sample_range] if features is not None else None
            fps_idx = sampler(sample_points_xyz.contiguous(), sample_features, npoint)
            indices.append(fps_idx + last)
            last += fps_sample_range
        return torch.cat(indices, dim=1)
