"""three_nn / three_interpolate -- the feature-propagation ops of the mmdet3d PointNet++ family.
Reference: mmdet3d/ops/interpolate/three_nn.py:9-46, three_interpolate.py:9-63 (forward only)."""
import torch

from ._common import OPS, _NoBackward, require


class ThreeNN(_NoBackward):
    @staticmethod
    def forward(ctx, target, source):
        """target (B, N, 3), source (B, M, 3) -> (dist (B, N, 3) L2 distances, idx int32 (B, N, 3)) of the 3 nearest sources."""
        require(target, "target")
        require(source, "source")
        B, N, _ = target.shape
        m = source.shape[1]
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=target.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=target.device)
        OPS.three_nn(B, N, m, target, source, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx


class ThreeInterpolate(_NoBackward):
    @staticmethod
    def forward(ctx, features, indices, weight):
        """features (B, C, M), indices int32 (B, n, 3), weight (B, n, 3) -> (B, C, n)."""
        require(features, "features")
        require(indices, "indices", torch.int32)
        require(weight, "weight")
        B, c, m = features.shape
        n = indices.shape[1]
        output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        OPS.three_interpolate(B, c, m, n, features, indices, weight, output)
        return output


three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply
