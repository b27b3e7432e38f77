"""Pointnet_Backbone -- the reference's 'Point Transformer' encoder (mmdet3d/models/backbone_net.py:27-124):
3 x (first-S sampling + kNN edge grouping + shared MLP + max + linear self-attention), 3 x attention
feature propagation, final 1x1 conv."""
from torch import nn

from .. import kernels as K
from ._packing import PackedModule, kmajor
from .pointnet2_utils import PointNetFeaturePropagationSA, PointNetSetAbstractionEdgeSA


class Pointnet_Backbone(PackedModule):
    def __init__(self, input_channels=3, use_xyz=True, conv_out=32, mul=1, radius=[0.3, 0.5, 0.7], nsample=[32, 48, 48]):
        super().__init__()
        sa1, sa2, sa3 = 32 * mul, 64 * mul, 128 * mul
        self.SA_modules = nn.ModuleList()
        for r, k, mlp in ((radius[0], nsample[0], [input_channels, sa1, sa1, sa1]),
                          (radius[1], nsample[1], [sa2, sa2, sa2, sa2]),
                          (radius[2], nsample[2], [sa3, sa3, sa3, sa3])):
            self.SA_modules.append(PointNetSetAbstractionEdgeSA(npoint=None, radius=r, nsample=k, mlp=mlp,
                                                                 sampling="RANDOM", use_xyz=use_xyz, use_knn=True))
        self.FP_modules = nn.ModuleList()
        self.FP_modules.append(PointNetFeaturePropagationSA(mlp=[67, sa1, sa1], mlp_inte=[sa2, 3, sa2, sa2, sa1]))
        self.FP_modules.append(PointNetFeaturePropagationSA(mlp=[160, sa3, sa2], mlp_inte=[sa3, sa1, sa3, sa2, sa2]))
        self.FP_modules.append(PointNetFeaturePropagationSA(mlp=[192, sa3, sa3], mlp_inte=[sa3, sa2, sa3, sa2, sa3]))
        self.cov_final = nn.Conv1d(sa1, conv_out, kernel_size=1)

    def _pack(self):
        return dict(w=kmajor(self.cov_final.weight), b=self.cov_final.bias.detach().float().contiguous())

    def _pack_key(self):   # only cov_final is packed here; children pack themselves
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in self.cov_final.parameters())

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def forward(self, pointcloud, numpoints):
        """pointcloud (B, N, 3 + input_channels), numpoints [S1, S2, S3] -> (xyz (B, N, 3), features (B, conv_out, N))."""
        xyz, features = self._break_up_pc(pointcloud.float())
        l_xyz, l_features = [xyz], [features]
        for i, sa in enumerate(self.SA_modules):
            li_xyz, li_features = sa(l_xyz[i], l_features[i], numpoints[i])
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        # the reference feeds xyz^T as the level-0 feature (backbone_net.py:117); read here point-major, no copy
        l_features[2] = self.FP_modules[2](l_xyz[2], l_xyz[3], l_features[2], l_features[3])
        l_features[1] = self.FP_modules[1](l_xyz[1], l_xyz[2], l_features[1], l_features[2])
        f0 = self.FP_modules[0](l_xyz[0], l_xyz[1], xyz, l_features[1], points1_point_major=True)
        pk = self.packed()
        return l_xyz[0], K.cn_linear(f0, pk["w"], bias=pk["b"])
