"""PointNet encoder (mmdet3d/models/pointnet.py:10-150): STN3d + STNkd(64) T-Nets, shared MLP 3-64-128-1024,
per-point features out (no global pooling at the end)."""
import torch
from torch import nn

from .. import kernels as K
from ._packing import PackedModule, fold_bn, kmajor


class _STN(PackedModule):
    def __init__(self, in_ch, kdim):
        super().__init__()
        self.conv1 = nn.Conv1d(in_ch, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, kdim * kdim)
        self.relu = nn.ReLU()
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)
        self._kdim = kdim

    def _pack(self):
        pk = {}
        for name, lin, bn in (("c1", self.conv1, self.bn1), ("c2", self.conv2, self.bn2), ("c3", self.conv3, self.bn3),
                              ("f1", self.fc1, self.bn4), ("f2", self.fc2, self.bn5)):
            w, b = fold_bn(lin.weight, lin.bias, bn)
            pk[name], pk[name + "b"] = w.t().contiguous(), b
        pk["f3"] = kmajor(self.fc3.weight)
        eye = torch.eye(self._kdim, device=self.fc3.bias.device, dtype=torch.float32).flatten()
        pk["f3b"] = (self.fc3.bias.detach().float() + eye).contiguous()     # "+ iden" (pointnet.py:38-44)
        return pk

    def forward(self, x):
        """x (B, C, N) -> per-object transform (B, kdim, kdim), row-major == k-major weight of x @ T."""
        self._inference_only()
        pk = self.packed()
        B = x.shape[0]
        h = K.cn_linear(x, pk["c1"], bias=pk["c1b"], act=K.ACT_RELU)
        h = K.cn_linear(h, pk["c2"], bias=pk["c2b"], act=K.ACT_RELU)
        h = K.cn_linear(h, pk["c3"], bias=pk["c3b"], act=K.ACT_RELU)
        g = K.cn_pool(h, mode=1, transposed=True)                       # (1, 1024, B): objects become the row axis
        g = K.cn_linear(g, pk["f1"], bias=pk["f1b"], act=K.ACT_RELU)
        g = K.cn_linear(g, pk["f2"], bias=pk["f2b"], act=K.ACT_RELU)
        t = K.cn_linear(g, pk["f3"], bias=pk["f3b"], y_pm=True)         # (1, B, kdim*kdim)
        return t.view(B, self._kdim, self._kdim)


class STN3d(_STN):
    def __init__(self, channel):
        super().__init__(channel, 3)


class STNkd(_STN):
    def __init__(self, k=64):
        super().__init__(k, k)
        self.k = k


class PointNetEncoder(PackedModule):
    def __init__(self, global_feat=True, feature_transform=False, channel=3):
        super().__init__()
        if channel != 3 or not feature_transform:
            raise NotImplementedError("the ReID configs build PointNet(normal_channel=False) with feature_transform=True")
        self.stn = STN3d(channel)
        self.conv1 = nn.Conv1d(channel, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.global_feat = global_feat
        self.feature_transform = feature_transform
        self.fstn = STNkd(k=64)

    def _pack_key(self):
        ts = [p for m in (self.conv1, self.conv2, self.conv3, self.bn1, self.bn2, self.bn3)
              for p in list(m.parameters()) + list(m.buffers())]
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in ts)

    def _pack(self):
        pk = {}
        for name, conv, bn in (("c1", self.conv1, self.bn1), ("c2", self.conv2, self.bn2), ("c3", self.conv3, self.bn3)):
            w, b = fold_bn(conv.weight, conv.bias, bn)
            pk[name], pk[name + "b"] = w.t().contiguous(), b
        return pk

    def forward(self, xyz):
        self._inference_only()
        pk = self.packed()
        x = xyz.float().contiguous()
        trans = self.stn(x)                                             # (B, 3, 3)
        x = K.cn_linear(x, trans)                                       # x^T @ trans, per-object weights
        x = K.cn_linear(x, pk["c1"], bias=pk["c1b"], act=K.ACT_RELU)
        trans_feat = self.fstn(x)                                       # (B, 64, 64)
        x = K.cn_linear(x, trans_feat)
        x = K.cn_linear(x, pk["c2"], bias=pk["c2b"], act=K.ACT_RELU)
        x = K.cn_linear(x, pk["c3"], bias=pk["c3b"])
        return xyz, x


class PointNet(nn.Module):
    def __init__(self, k=40, normal_channel=True):
        super().__init__()
        channel = 6 if normal_channel else 3
        self.feat = PointNetEncoder(global_feat=True, feature_transform=True, channel=channel)

    def forward(self, x, backbone_list=None):
        return self.feat(x)
