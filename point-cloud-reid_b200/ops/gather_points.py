"""gather_points(features (B,C,N), indices int32 (B,M)) -> (B,C,M).
Reference: mmdet3d/ops/gather_points/gather_points.py:7-50 (forward only)."""
import torch

from ._common import OPS, _NoBackward, require


class GatherPoints(_NoBackward):
    @staticmethod
    def forward(ctx, features, indices):
        require(features, "features")
        require(indices, "indices", torch.int32)
        B, npoint = indices.shape
        _, C, N = features.shape
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        OPS.gather_points(B, C, N, npoint, features, indices, output)
        ctx.mark_non_differentiable(indices)
        return output


gather_points = GatherPoints.apply
