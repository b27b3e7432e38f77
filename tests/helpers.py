"""Shared test helpers: the reference's model configs (configs_reid/_base_/reidentifiers/reid_pts_*.py),
model + oracle builders with identical weights, golden loader."""
import os

import numpy as np
import torch

from oracle import reid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_DS = [dict(type='LinearRes', n_in=1024, n_out=512, norm='GN', ng=64), dict(type='LinearRes', n_in=512, n_out=128, norm='GN', ng=16),
       dict(type='Linear', in_features=128, out_features=64)]


def model_cfg(kind="pt", backbone_list=(128, 64, 32)):
    """kind: pt (reid_pts_point-transformer_point-cat.py), concat (reid_pts_point-transformer_baseline.py),
    dgcnn (reid_pts_dgcnn_point-cat.py), pointnet (reid_pts_pointnet_point-cat.py)."""
    c = dict(type='ReIDNet', hidden_size=128, combine='point-cat', match_type='xcorr_eff', pool_type='both',
             backbone_list=list(backbone_list), output_sequence_size=64,
             backbone=dict(type='Pointnet_Backbone', input_channels=0, use_xyz=True, conv_out=64),
             match_head=[dict(type='LinearRes', n_in=128, n_out=128, norm='GN', ng=8), dict(type='Linear', in_features=128, out_features=1)],
             downsample=None, cls_head=None, fp_head=None, shape_head=None,
             cross_stage1=dict(type='corss_attention', d_model=64, nhead=2, attention='linear'),
             cross_stage2=dict(type='corss_attention', d_model=64, nhead=2, attention='linear'),
             local_stage1=dict(), local_stage2=dict())
    if kind == "concat":
        c.update(match_type='concat', combine='cat', pool_type='max', cross_stage1=None, cross_stage2=None, local_stage1=None,
                 local_stage2=None, match_head=[dict(type='LinearRes', n_in=256, n_out=256, norm='GN', ng=32),
                                                dict(type='Linear', in_features=256, out_features=1)])
    elif kind in ("xcorr", "xcorr-baseline"):
        # reid_pts_point-transformer_baseline_orig.py ('xcorr': local_self_attention stages, knum 48) /
        # reid_pts_point-transformer_baseline_stnet.py ('xcorr-baseline')
        c.update(match_type=kind)
        if kind == "xcorr":
            loc = dict(type='local_self_attention', d_model=64, nhead=2, attention='linear', knum=48, pos_size=64)
            c.update(local_stage1=dict(loc), local_stage2=dict(loc))
    elif kind in ("pt15m", "pt7m"):
        # reid_pts_point-transformer-1.5M_point-cat.py (mul=2, width 64) / -7M_point-cat.py (mul=4, width 128)
        mul, w, ng = (2, 64, 8) if kind == "pt15m" else (4, 128, 16)
        c.update(hidden_size=2 * w, output_sequence_size=w,
                 backbone=dict(type='Pointnet_Backbone', input_channels=0, use_xyz=True, conv_out=w, mul=mul),
                 match_head=[dict(type='LinearRes', n_in=2 * w, n_out=2 * w, norm='GN', ng=ng),
                             dict(type='Linear', in_features=2 * w, out_features=1)],
                 cross_stage1=dict(type='corss_attention', d_model=w, nhead=2, attention='linear'),
                 cross_stage2=dict(type='corss_attention', d_model=w, nhead=2, attention='linear'))
    elif kind == "dgcnn":
        c.update(use_dgcnn=True, backbone=dict(type='dgcnn', dropout=0.5, emb_dims=1024, k=20, output_channels=40), downsample=_DS,
                 match_head=[dict(type='LinearRes', n_in=128, n_out=128, norm='GN', ng=16), dict(type='Linear', in_features=128, out_features=1)])
    elif kind == "pointnet":
        c.update(use_dgcnn=True, backbone=dict(type='PointNet', k=40, normal_channel=False), downsample=_DS)
    return c


ORACLE_KW = {
    "pt": dict(backbone='Pointnet_Backbone'),
    "concat": dict(backbone='Pointnet_Backbone', match_type='concat', pool_type='max', combine='cat', head_ng=32),
    "dgcnn": dict(backbone='dgcnn', head_ng=16),
    "pointnet": dict(backbone='PointNet', head_ng=8),
    "xcorr": dict(backbone='Pointnet_Backbone', match_type='xcorr', knum=48),
    "xcorr-baseline": dict(backbone='Pointnet_Backbone', match_type='xcorr-baseline'),
    "pt15m": dict(backbone='Pointnet_Backbone', head_ng=8),
    "pt7m": dict(backbone='Pointnet_Backbone', head_ng=16),
}


def build_pair(kind="pt", backbone_list=(128, 64, 32), device="cpu", perturb=True):
    """-> (product model on `device`, oracle) sharing seed-66 weights (+ deterministic norm perturbation)."""
    from pcreid_b200.models import build_model
    torch.manual_seed(66)
    m = build_model(model_cfg(kind, backbone_list)).eval()
    sd = m.state_dict()
    if perturb:
        sd = O.perturb_norm_state(sd)
        m.load_state_dict(sd)
    orc = O.ReIDOracle(sd, backbone_list=backbone_list, **ORACLE_KW[kind])
    return m.to(device), orc


def image_cfg(dim=192, downsample_dim=64):
    """reid_image_deit-tiny_point-cat.py (dim 192) / reid_image_deit-base_point-cat.py (dim 768); aux heads as shipped."""
    hp = dim * 2
    aux = lambda n_out: [dict(type='LinearRes', n_in=hp, n_out=hp, norm='GN', ng=64), dict(type='Linear', in_features=hp, out_features=n_out)]
    return dict(type='ImageReIDNet', dim=dim, backbone='deit-tiny', downsample_dim=downsample_dim, combine='point-cat',
                match_type='xcorr_eff', pool_type='both',
                losses_to_use=dict(kl=False, match=True, cls=True, fp=True, triplet=False, vis=True),
                downsample=[dict(type='LinearRes', n_in=dim, n_out=256, norm='GN', ng=32), dict(type='LinearRes', n_in=256, n_out=128, norm='GN', ng=16),
                            dict(type='Linear', in_features=128, out_features=downsample_dim)],
                cross_lin_attn=dict(type='cross_lin_attn', d_model=downsample_dim, nhead=2, attention='linear'),
                cls_head=aux(20), fp_head=aux(1), vis_head=aux(4),
                match_head=[dict(type='LinearRes', n_in=2 * downsample_dim, n_out=2 * downsample_dim, norm='GN', ng=16),
                            dict(type='Linear', in_features=2 * downsample_dim, out_features=1)])


def build_image_pair(dim=192, downsample_dim=64, device="cpu", perturb=True):
    """-> (ImageReIDNet on `device`, oracle.ImageReIDOracle) sharing seed-66 weights."""
    from pcreid_b200.models import build_model
    torch.manual_seed(66)
    m = build_model(image_cfg(dim, downsample_dim)).eval()
    sd = m.state_dict()
    if perturb:
        sd = O.perturb_norm_state(sd)
        m.load_state_dict(sd)
    orc = O.ImageReIDOracle(sd, downsample_dim=downsample_dim, downsample_ng=(32, 16), head_ng=16)
    return m.to(device), orc


def weight_checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def margin_aware_top1(ref, got, tol):
    """rows of `ref` whose top-1/top-2 gap exceeds 2*tol must keep their arg-max (SURVEY.md 8d parity gates)."""
    top2 = torch.topk(ref, 2, dim=1)[0]
    decisive = (top2[:, 0] - top2[:, 1]) > 2 * tol
    same = ref.argmax(1) == got.argmax(1)
    return bool(same[decisive].all()), float(same.float().mean()), int(decisive.sum())
