"""ball_query(min_radius, max_radius, sample_num, xyz, center_xyz) -> int32 (B, npoint, sample_num).
Reference: mmdet3d/ops/ball_query/ball_query.py:7-54."""
import torch

from ._common import OPS, require


class BallQuery:
    @staticmethod
    def apply(min_radius, max_radius, sample_num, xyz, center_xyz):
        require(center_xyz, "center_xyz")
        require(xyz, "xyz")
        assert min_radius < max_radius
        B, N, _ = xyz.shape
        npoint = center_xyz.shape[1]
        idx = torch.zeros((B, npoint, sample_num), dtype=torch.int32, device=xyz.device)
        OPS.ball_query(B, N, npoint, float(min_radius), float(max_radius), sample_num, center_xyz, xyz, idx)
        return idx

    forward = apply


ball_query = BallQuery.apply
