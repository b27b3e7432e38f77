"""mmdet3d PointNet++ module family on top of the point ops (inference path):
PointSAModuleMSG / PointSAModule (mmdet3d/ops/pointnet_modules/point_sa_module.py:8-354) and PointFPModule
(point_fp_module.py:10-79), with a minimal stand-in for mmcv's ConvModule (Conv2d 1x1 -> BatchNorm2d -> ReLU,
state_dict keys `conv.weight[, conv.bias]`, `bn.*`) so that mmdet3d checkpoints of these modules load unchanged.

Execution differs from the reference, results do not: the grouped (B, C+3, S, k) tensor is never materialised.  The
first 1x1 conv is linear in [xyz_j - centre, feat_j], so it is split into a per-point term W.[xyz_j ; feat_j] and a
per-centre term -Wa.centre (+ folded BatchNorm shift); the edge tensor is built after that conv
(pcreid_edge_build) or, for three equal-width layers, the whole MLP + max runs in the fused pcreid_sa_edge_mlp kernel.
"""
import torch
from torch import nn

from .. import kernels as K
from ..models._packing import PackedModule, fold_bn
from .ball_query import ball_query
from .furthest_point_sample import Points_Sampler
from .gather_points import gather_points
from .interpolate import three_interpolate, three_nn
from ._common import OPS


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule subset used by the PointNet++ modules: 1x1 Conv2d, optional BN2d, ReLU."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), conv_cfg=dict(type="Conv2d"),
                 norm_cfg=dict(type="BN2d"), act_cfg=dict(type="ReLU"), bias="auto"):
        super().__init__()
        if tuple(kernel_size) != (1, 1) or tuple(stride) != (1, 1) or (conv_cfg or {}).get("type", "Conv2d") != "Conv2d":
            raise NotImplementedError("ConvModule stand-in: 1x1 Conv2d only")
        if norm_cfg is not None and norm_cfg.get("type") not in ("BN2d", "BN"):
            raise NotImplementedError(f"ConvModule stand-in: norm {norm_cfg} not supported")
        if act_cfg is not None and act_cfg.get("type") != "ReLU":
            raise NotImplementedError(f"ConvModule stand-in: activation {act_cfg} not supported")
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, 1, bias=bias)
        if self.with_norm:
            self.bn = nn.BatchNorm2d(out_channels)
        if self.with_activation:
            self.activate = nn.ReLU(inplace=True)

    def folded(self):
        """-> (W (C_out, C_in), b (C_out)) with eval BatchNorm folded, act code."""
        if self.with_norm:
            w, b = fold_bn(self.conv.weight, self.conv.bias, self.bn)
        else:
            w = self.conv.weight.detach().reshape(self.conv.weight.shape[0], -1).float()
            b = self.conv.bias.detach().float() if self.conv.bias is not None else w.new_zeros(w.shape[0])
        return w.contiguous(), b.contiguous(), (K.ACT_RELU if self.with_activation else K.ACT_NONE)


def _mlp_tail(x, layers):
    """remaining ConvModules over a channel-major (B, C, R) tensor."""
    for w, b, act in layers:
        x = K.cn_linear(x, w.t().contiguous(), bias=b, act=act)
    return x


class _PackedMLPs(PackedModule):
    def _inference(self):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: pcreid_b200 implements the inference path only; call .eval()")


class BasePointSAModule(_PackedMLPs):
    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=["D-FPS"], fps_sample_range_list=[-1],
                 dilated_group=False, use_xyz=True, pool_mod="max", normalize_xyz=False,
                 grouper_return_grouped_xyz=False, grouper_return_grouped_idx=False):
        super().__init__()
        assert len(radii) == len(sample_nums) == len(mlp_channels)
        assert pool_mod in ["max", "avg"]
        assert isinstance(fps_mod, (list, tuple)) and isinstance(fps_sample_range_list, (list, tuple))
        assert len(fps_mod) == len(fps_sample_range_list)
        if isinstance(mlp_channels, tuple):
            mlp_channels = list(map(list, mlp_channels))
        self.mlp_channels = [list(m) for m in mlp_channels]
        if isinstance(num_point, int):
            self.num_point = [num_point]
        elif isinstance(num_point, (list, tuple)) or num_point is None:
            self.num_point = num_point
        else:
            raise NotImplementedError("Error type of num_point!")
        self.pool_mod = pool_mod
        self.use_xyz = use_xyz
        self.normalize_xyz = normalize_xyz
        self.radii, self.sample_nums = list(radii), list(sample_nums)
        self.min_radii = [(radii[i - 1] if (dilated_group and i != 0) else 0) for i in range(len(radii))]
        self.mlps = nn.ModuleList()
        self.fps_mod_list = fps_mod
        self.fps_sample_range_list = fps_sample_range_list
        if self.num_point is not None:
            self.points_sampler = Points_Sampler(self.num_point, self.fps_mod_list, self.fps_sample_range_list)

    def _sample_points(self, points_xyz, features, indices, target_xyz):
        xyz_flipped = points_xyz.transpose(1, 2).contiguous()
        if indices is not None:
            assert indices.shape[1] == self.num_point[0]
            new_xyz = gather_points(xyz_flipped, indices).transpose(1, 2).contiguous() if self.num_point is not None else None
        elif target_xyz is not None:
            new_xyz = target_xyz.contiguous()
        elif self.num_point is not None:
            indices = self.points_sampler(points_xyz, features)
            new_xyz = gather_points(xyz_flipped, indices).transpose(1, 2).contiguous()
        else:
            new_xyz = None
        return new_xyz, indices

    def _pack(self):
        packs = []
        for i, mlp in enumerate(self.mlps):
            layers = [m.folded() for m in mlp]
            w1, b1, act1 = layers[0]
            if self.use_xyz:
                wa, wf = w1[:, :3], w1[:, 3:]
                if self.normalize_xyz and self.num_point is not None:
                    wa = wa / self.radii[i]
            else:
                wa, wf = None, w1
            packs.append(dict(wa=None if wa is None else wa.t().contiguous(), nwa=None if wa is None else (-wa).t().contiguous(),
                              wf=wf.t().contiguous() if wf.shape[1] > 0 else None, b1=b1, act1=act1, tail=layers[1:]))
        return dict(scales=packs)

    def forward(self, points_xyz, features=None, indices=None, target_xyz=None):
        """points_xyz (B, N, 3), features (B, C, N) -> new_xyz (B, M, 3), new_features (B, sum_k mlp[k][-1], M), indices."""
        self._inference()
        points_xyz = points_xyz.contiguous().float()
        new_xyz, indices = self._sample_points(points_xyz, features, indices, target_xyz)
        pk = self.packed()["scales"]
        outs = []
        for i, sc in enumerate(pk):
            if sc["act1"] != K.ACT_RELU:
                raise NotImplementedError("first ConvModule without ReLU")
            if new_xyz is None:                                  # GroupAll: one group holding every point, absolute xyz
                h = self._first_conv_points(points_xyz, features, sc, bias=sc["b1"], act=K.ACT_RELU)
                h = _mlp_tail(h, sc["tail"])
                pooled = K.cn_pool(h, mode=1) if self.pool_mod == "max" else K.cn_pool(h, mode=0)[:, h.shape[1]:].contiguous()
                outs.append(pooled.unsqueeze(-1))
                continue
            idx = ball_query(self.min_radii[i], self.radii[i], self.sample_nums[i], points_xyz, new_xyz)
            p1 = self._first_conv_points(points_xyz, features, sc)                        # (B, C1, N)
            if sc["nwa"] is not None:
                cc = K.cn_linear(new_xyz, sc["nwa"], bias=sc["b1"], x1_pm=True)           # (B, C1, S)
            else:
                cc = sc["b1"].view(1, -1, 1).expand(p1.shape[0], -1, new_xyz.shape[1]).contiguous()
            outs.append(self._edge_mlp_pool(p1, cc, idx, sc["tail"], self.pool_mod))
        return new_xyz, torch.cat(outs, dim=1), indices

    @staticmethod
    def _first_conv_points(points_xyz, features, sc, bias=None, act=K.ACT_NONE):
        if sc["wa"] is not None and sc["wf"] is not None and features is not None:
            return K.cn_linear(points_xyz, sc["wa"], x2=features.contiguous().float(), w2=sc["wf"], bias=bias, act=act, x1_pm=True)
        if sc["wa"] is not None:
            assert features is None, "features given but the first conv has no feature channels"
            return K.cn_linear(points_xyz, sc["wa"], bias=bias, act=act, x1_pm=True)
        return K.cn_linear(features.contiguous().float(), sc["wf"], bias=bias, act=act)

    @staticmethod
    def _edge_mlp_pool(p1, cc, idx, tail, pool_mod="max"):
        """relu(P1[idx] + Cc) -> remaining ConvModules -> max (or mean, point_sa_module.py:144-164) over the k samples."""
        B, C, N = p1.shape
        S, k = idx.shape[1], idx.shape[2]
        same = len(tail) == 2 and all(w.shape == (C, C) and act == K.ACT_RELU for w, _, act in tail)
        if pool_mod == "max" and same and C % 16 == 0 and C <= 128 and k <= 128:
            (w2, b2, _), (w3, b3, _) = tail
            return K.sa_edge_mlp(p1, cc, idx, w2.t().contiguous(), b2, w3.t().contiguous(), b3)
        Co = tail[-1][0].shape[0] if tail else C
        out = torch.empty((B, Co, S), device=p1.device, dtype=torch.float32)
        widest = max([C] + [w.shape[0] for w, _, _ in tail])
        step = max(1, (256 << 20) // (widest * S * k * 4))
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            nb = b1 - b0
            h = torch.empty((nb, C, S * k), device=p1.device, dtype=torch.float32)
            OPS.edge_build(nb, C, N, S, k, p1[b0:b1], cc[b0:b1], idx[b0:b1], h)
            h = _mlp_tail(h, tail)
            (OPS.seg_max if pool_mod == "max" else OPS.seg_mean)(nb * Co * S, k, h, out[b0:b1])
        return out


class PointSAModuleMSG(BasePointSAModule):
    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=["D-FPS"], fps_sample_range_list=[-1],
                 dilated_group=False, norm_cfg=dict(type="BN2d"), use_xyz=True, pool_mod="max", normalize_xyz=False,
                 bias="auto"):
        super().__init__(num_point=num_point, radii=radii, sample_nums=sample_nums, mlp_channels=mlp_channels, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list, dilated_group=dilated_group, use_xyz=use_xyz,
                         pool_mod=pool_mod, normalize_xyz=normalize_xyz)
        for i in range(len(self.mlp_channels)):
            mlp_channel = self.mlp_channels[i]
            if use_xyz:
                mlp_channel[0] += 3
            mlp = nn.Sequential()
            for j in range(len(mlp_channel) - 1):
                mlp.add_module(f"layer{j}", ConvModule(mlp_channel[j], mlp_channel[j + 1], kernel_size=(1, 1), stride=(1, 1),
                                                       conv_cfg=dict(type="Conv2d"), norm_cfg=norm_cfg, bias=bias))
            self.mlps.append(mlp)


class PointSAModule(PointSAModuleMSG):
    def __init__(self, mlp_channels, num_point=None, radius=None, num_sample=None, norm_cfg=dict(type="BN2d"), use_xyz=True,
                 pool_mod="max", fps_mod=["D-FPS"], fps_sample_range_list=[-1], normalize_xyz=False):
        super().__init__(mlp_channels=[mlp_channels], num_point=num_point, radii=[radius], sample_nums=[num_sample],
                         norm_cfg=norm_cfg, use_xyz=use_xyz, pool_mod=pool_mod, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list, normalize_xyz=normalize_xyz)


SA_MODULES = {"PointSAModule": PointSAModule, "PointSAModuleMSG": PointSAModuleMSG}


def build_sa_module(cfg, *args, **kwargs):
    """pointnet_modules/builder.py:6-38."""
    if cfg is None:
        cfg_ = dict(type="PointSAModule")
    else:
        if not isinstance(cfg, dict):
            raise TypeError("cfg must be a dict")
        if "type" not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = cfg.copy()
    module_type = cfg_.pop("type")
    if module_type not in SA_MODULES:
        raise KeyError(f"Unrecognized module type {module_type}")
    return SA_MODULES[module_type](*args, **kwargs, **cfg_)


class PointFPModule(_PackedMLPs):
    """three_nn inverse-distance interpolation of the source features onto the target points, concat, shared MLP."""

    def __init__(self, mlp_channels, norm_cfg=dict(type="BN2d"), init_cfg=None):
        super().__init__()
        self.fp16_enabled = False
        self.mlps = nn.Sequential()
        for i in range(len(mlp_channels) - 1):
            self.mlps.add_module(f"layer{i}", ConvModule(mlp_channels[i], mlp_channels[i + 1], kernel_size=(1, 1), stride=(1, 1),
                                                         conv_cfg=dict(type="Conv2d"), norm_cfg=norm_cfg))

    def _pack(self):
        return dict(layers=[m.folded() for m in self.mlps])

    def forward(self, target, source, target_feats, source_feats):
        """target (B, n, 3), source (B, m, 3), target_feats (B, C1, n) | None, source_feats (B, C2, m) -> (B, mlp[-1], n)."""
        self._inference()
        if source is not None:
            dist, idx = three_nn(target.contiguous().float(), source.contiguous().float())
            dist_reciprocal = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_reciprocal, dim=2, keepdim=True)
            weight = dist_reciprocal / norm
            interpolated = three_interpolate(source_feats.contiguous().float(), idx, weight.contiguous())
        else:
            interpolated = source_feats.expand(*source_feats.size()[0:2], target.size(1)).contiguous()
        layers = self.packed()["layers"]
        w1, b1, act1 = layers[0]
        if target_feats is not None:                 # cat([interpolated, target_feats]) folded into a two-operand first conv
            c2 = interpolated.shape[1]
            x = K.cn_linear(interpolated, w1[:, :c2].t().contiguous(), x2=target_feats.contiguous().float(),
                            w2=w1[:, c2:].t().contiguous(), bias=b1, act=act1)
        else:
            x = K.cn_linear(interpolated, w1.t().contiguous(), bias=b1, act=act1)
        return _mlp_tail(x, layers[1:])
