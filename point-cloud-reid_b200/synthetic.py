"""Synthetic workloads of the named shapes (BASELINE.json configs) and the shipped model configs the
benchmarks instantiate with random-init weights.  Product-side copy: bench.py's measured legs and scripts use this module,
never oracle/ (tests/test_host_logic.py asserts that both generators produce identical tensors)."""
import torch


def synth_objects(B, N, seed, dup=False):
    """fp32 xyz (B, N, 3) = randn * (2.0, 0.9, 0.8) -- a vehicle-sized anisotropic blob per object; ``dup`` mimics the
    resample-with-replacement of subsamplePC (mmdet3d/datasets/utils.py:606-621): max(2, N // 6) unique points."""
    g = torch.Generator().manual_seed(seed)
    scale = torch.tensor([2.0, 0.9, 0.8])
    if not dup:
        return torch.randn(B, N, 3, generator=g) * scale
    U = max(2, N // 6)
    base = torch.randn(B, U, 3, generator=g) * scale
    pick = torch.randint(0, U, (B, N), generator=g)
    return torch.gather(base, 1, pick.unsqueeze(-1).expand(B, N, 3)).contiguous()


def synth_tokens(B, C, S, seed):
    """fp32 token maps (B, C, S) ~ N(0, 1): stand-in for the last hidden state of an image backbone."""
    return torch.randn(B, C, S, generator=torch.Generator().manual_seed(seed))


def point_transformer_cfg(backbone_list=(256, 128, 64)):
    """configs_reid/_base_/reidentifiers/reid_pts_point-transformer_point-cat.py with the backbone_list of the
    num_point ablation configs (pts_point-transformer_r_nus_det_400e_256pts.py:7-9)."""
    xa = dict(type='corss_attention', d_model=64, nhead=2, attention='linear')
    return dict(type='ReIDNet', hidden_size=128, combine='point-cat', match_type='xcorr_eff', pool_type='both',
                backbone_list=list(backbone_list), output_sequence_size=64,
                backbone=dict(type='Pointnet_Backbone', input_channels=0, use_xyz=True, conv_out=64),
                match_head=[dict(type='LinearRes', n_in=128, n_out=128, norm='GN', ng=8),
                            dict(type='Linear', in_features=128, out_features=1)],
                downsample=None, cls_head=None, fp_head=None, shape_head=None, cross_stage1=dict(xa), cross_stage2=dict(xa),
                local_stage1=dict(), local_stage2=dict())


_DOWNSAMPLE = [dict(type='LinearRes', n_in=1024, n_out=512, norm='GN', ng=64), dict(type='LinearRes', n_in=512, n_out=128, norm='GN', ng=16),
               dict(type='Linear', in_features=128, out_features=64)]


def dgcnn_cfg(backbone_list=(256, 128, 64)):
    """configs_reid/_base_/reidentifiers/reid_pts_dgcnn_point-cat.py: DGCNN (k = 20, emb_dims 1024) + per-point downsample to 64."""
    c = point_transformer_cfg(backbone_list)
    c.update(use_dgcnn=True, backbone=dict(type='dgcnn', dropout=0.5, emb_dims=1024, k=20, output_channels=40),
             downsample=[dict(d) for d in _DOWNSAMPLE],
             match_head=[dict(type='LinearRes', n_in=128, n_out=128, norm='GN', ng=16), dict(type='Linear', in_features=128, out_features=1)])
    return c


def pointnet_cfg(backbone_list=(128, 64, 32)):
    """configs_reid/_base_/reidentifiers/reid_pts_pointnet_point-cat.py: PointNet (T-Nets, 1024-wide trunk) + downsample to 64."""
    c = point_transformer_cfg(backbone_list)
    c.update(use_dgcnn=True, backbone=dict(type='PointNet', k=40, normal_channel=False), downsample=[dict(d) for d in _DOWNSAMPLE])
    return c
