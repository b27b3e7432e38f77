"""knn(k, xyz, center_xyz=None, transposed=False) -> int32 (B, k, npoint).
Reference: mmdet3d/ops/knn/knn.py:7-71 (heap kNN, k <= 100)."""
import torch

from ._common import OPS, require


class KNN:
    @staticmethod
    def apply(k, xyz, center_xyz=None, transposed=False, return_dist=False):
        assert k > 0
        if center_xyz is None:
            center_xyz = xyz
        if transposed:
            xyz = xyz.transpose(2, 1).contiguous()
            center_xyz = center_xyz.transpose(2, 1).contiguous()
        require(xyz, "xyz")
        require(center_xyz, "center_xyz")
        assert xyz.get_device() == center_xyz.get_device(), "center_xyz and xyz should be put on the same device"
        assert k <= 100, "k should be no larger than 100 (knn.py:30)"
        B, npoint, _ = center_xyz.shape
        N = xyz.shape[1]
        idx = torch.zeros((B, k, npoint), dtype=torch.int32, device=xyz.device)
        dist2 = torch.zeros((B, k, npoint), dtype=torch.float32, device=xyz.device)
        OPS.knn_t(B, N, npoint, k, xyz, center_xyz, idx, dist2)
        return (idx, dist2) if return_dist else idx

    forward = apply


knn = KNN.apply
