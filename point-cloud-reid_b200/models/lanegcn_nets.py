"""LinearRes (mmdet3d/models/lanegcn_nets.py:193-241): Linear-GN-ReLU-Linear-GN (+ transform) + residual, ReLU.
Used by the match head (rows = pairs) and by the DGCNN / PointNet `downsample` (rows = points)."""
from math import gcd

import torch
from torch import nn

from .. import kernels as K
from ._packing import PackedModule, kmajor


class LinearRes(PackedModule):
    def __init__(self, n_in, n_out, norm='GN', ng=32, activation='ReLU'):
        super().__init__()
        assert norm in ['GN', 'BN', 'SyncBN']
        if norm != 'GN' or activation != 'ReLU':
            raise NotImplementedError("only norm='GN', activation='ReLU' is used by the ReID configs")
        self.linear1 = nn.Linear(n_in, n_out, bias=False)
        self.linear2 = nn.Linear(n_out, n_out, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.groups = gcd(ng, n_out)
        self.norm1 = nn.GroupNorm(self.groups, n_out)
        self.norm2 = nn.GroupNorm(self.groups, n_out)
        if n_in != n_out:
            self.transform = nn.Sequential(nn.Linear(n_in, n_out, bias=False), nn.GroupNorm(self.groups, n_out))
        else:
            self.transform = None

    def _pack(self):
        f = lambda t: t.detach().float().contiguous()
        pk = dict(w1=kmajor(self.linear1.weight), w2=kmajor(self.linear2.weight),
                  g1=f(self.norm1.weight), b1=f(self.norm1.bias), g2=f(self.norm2.weight), b2=f(self.norm2.bias))
        if self.transform is not None:
            pk.update(wt=kmajor(self.transform[0].weight), gt=f(self.transform[1].weight), bt=f(self.transform[1].bias))
        return pk

    def forward_cn(self, x, x_pm=False):
        """channel-major core: x (B, n_in, R) -> (B, n_out, R); every column (point / pair) is one row of the reference.
        x_pm: x is point-major (B, R, n_in) -- the reference's own row layout -- and is read in place."""
        pk = self.packed()
        h = K.cn_groupnorm(K.cn_linear(x, pk["w1"], x1_pm=x_pm), pk["g1"], pk["b1"], self.groups, act=K.ACT_RELU)
        h = K.cn_linear(h, pk["w2"])
        if self.transform is not None:
            t = K.cn_groupnorm(K.cn_linear(x, pk["wt"], x1_pm=x_pm), pk["gt"], pk["bt"], self.groups)
            return K.cn_groupnorm(h, pk["g2"], pk["b2"], self.groups, res=t, act=K.ACT_RELU)
        if x_pm:
            x = x.transpose(1, 2).contiguous()          # the identity shortcut is read channel-major
        return K.cn_groupnorm(h, pk["g2"], pk["b2"], self.groups, res=x, act=K.ACT_RELU)

    def forward(self, x):
        """reference API: x (R, n_in) -> (R, n_out)."""
        xc = x.float().t().contiguous().unsqueeze(0)
        return self.forward_cn(xc)[0].t().contiguous()
