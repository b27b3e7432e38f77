"""STNet-style PointNet++ blocks of the reference 'Point Transformer' backbone, running on the
pcreid CUDA kernels.  Class names, constructor arguments and state_dict keys follow
mmdet3d/models/pointnet2_utils.py (LinearAttention :14-47, Self_Attention :55-114,
PointNetSetAbstractionEdgeSA :290-360, FP_SA :362-437, PointNetFeaturePropagationSA :439-473)."""
import torch
from torch import nn

from .. import kernels as K
from ._packing import PackedModule, fold_bn, kmajor


class LinearAttention(nn.Module):
    """Parameter-free; kept so that module trees (and repr) match the reference.  The arithmetic
    (elu+1 feature map, KV summary, 1/(Q.Ksum+eps)) lives in pcreid_linattn_kv / pcreid_linattn_scale."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps


def fused_attention_supported(d, nhead, f1, f2, out):
    """shapes csrc/attn_tc.cu is built for (the mul=1 configurations); wider models stay on the unfused kernels."""
    return (d % 16 == 0 and d <= 128 and nhead <= 4 and d % nhead == 0 and (d // nhead) % 16 == 0 and f2 % 16 == 0 and f2 <= 128
            and (f1 + 7) // 8 * 8 <= 128 and out % 16 == 0 and out <= 2 * d)


def attention_message(q, wkv, ksum, nhead, s_len, pk, q_map=None, t_map=None, B=None):
    """scale -> (Q.KV) -> merge -> LayerNorm1.  The per-object KV summary and the merge projection are multiplied once per
    template object (d x d x d), so each point sees ONE d x d GEMM instead of two."""
    qs = K.linattn_scale(q, ksum, nhead, s_len, q_map=q_map, ksum_map=t_map, B=B)
    m = K.cn_linear(wkv, pk["merge"], x1_pm=True, y_pm=True)        # (B_t, d, d) = blockdiag(KV) Wm^T, k-major per object
    msg = K.cn_linear(qs, m, w1_map=t_map)
    return K.cn_groupnorm(msg, pk["n1w"], pk["n1b"], 1)


def attention_ffn(feat, msg, pk, residual, feat_map=None, feat_pm=False):
    """mlp(cat[feat, msg]) -> LayerNorm2 (-> + feat)."""
    hid = K.cn_linear(feat, pk["mlp0a"], x2=msg, w2=pk["mlp0b"], act=K.ACT_RELU, x1_map=feat_map, x1_pm=feat_pm,
                      B=msg.shape[0])
    m2 = K.cn_linear(hid, pk["mlp2"])
    return K.cn_groupnorm(m2, pk["n2w"], pk["n2b"], 1, res=feat if residual else None, r_map=feat_map)


class Self_Attention(PackedModule):
    def __init__(self, d_model, nhead, attention='linear'):
        super().__init__()
        self.dim = d_model // nhead
        self.nhead = nhead
        self.pos_mlp = nn.Sequential(nn.Linear(3, d_model), nn.ReLU(), nn.Linear(d_model, d_model))
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.attention = LinearAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(nn.Linear(d_model * 2, d_model * 2, bias=False), nn.ReLU(True),
                                 nn.Linear(d_model * 2, d_model, bias=False))
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.tc_mode = False      # True: the block runs as attn_front / linattn_kv_img / attn_back on the tensor cores (fast mode)

    def _pack(self):
        d = self.q_proj.weight.shape[0]
        m0 = self.mlp[0].weight.detach()
        pk = dict(
            pos0=kmajor(self.pos_mlp[0].weight), pos0b=self.pos_mlp[0].bias.detach().float().contiguous(),
            pos2=kmajor(self.pos_mlp[2].weight), pos2b=self.pos_mlp[2].bias.detach().float().contiguous(),
            qkv=kmajor(torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], 0)),
            merge=kmajor(self.merge.weight), mlp0a=kmajor(m0[:, :d]), mlp0b=kmajor(m0[:, d:]), mlp2=kmajor(self.mlp[2].weight),
            n1w=self.norm1.weight.detach().float().contiguous(), n1b=self.norm1.bias.detach().float().contiguous(),
            n2w=self.norm2.weight.detach().float().contiguous(), n2b=self.norm2.bias.detach().float().contiguous())
        if fused_attention_supported(d, self.nhead, d, d, d):
            pk["front_blob"] = K.weight_blob(K.tf32_image(self.pos_mlp[2].weight),
                                             K.tf32_image(torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], 0)))
            pk["back_blob"] = K.weight_blob(K.tf32_image(self.merge.weight), K.tf32_image(m0[:, :d]), K.tf32_image(m0[:, d:]),
                                            K.tf32_image(self.mlp[2].weight))
        return pk

    def forward(self, feat, xyz, mask=None):
        """feat (B, C, N), xyz (B, N, 3) -> (B, C, N)."""
        assert mask is None, "masks are never used on the ReID path"
        pk = self.packed()
        C, S = feat.shape[1], feat.shape[2]
        if self.tc_mode and "front_blob" in pk:
            qkv = K.attn_front(xyz, feat, pk["pos0"], pk["pos0b"], pk["pos2b"], pk["front_blob"], C, 3 * C, 0)
            kvimg, ksum = K.linattn_kv_img(qkv[:, C:2 * C], qkv[:, 2 * C:], self.nhead, S)
            return K.attn_back(feat, qkv[:, :C], ksum, kvimg, pk["n1w"], pk["n1b"], pk["n2w"], pk["n2b"], pk["back_blob"],
                               self.nhead, C, residual=True)
        hid = K.cn_linear(xyz, pk["pos0"], bias=pk["pos0b"], act=K.ACT_RELU, x1_pm=True)
        feat_pos = K.cn_linear(hid, pk["pos2"], bias=pk["pos2b"], res=feat)
        qkv = K.cn_linear(feat_pos, pk["qkv"])
        wkv, ksum = K.linattn_kv(qkv[:, C:2 * C], qkv[:, 2 * C:], self.nhead)
        msg = attention_message(qkv[:, :C], wkv, ksum, self.nhead, S, pk)
        return attention_ffn(feat, msg, pk, residual=True)


class PointNetSetAbstractionEdgeSA(PackedModule):
    def __init__(self, npoint, radius, nsample, mlp, sampling, use_xyz=True, group_all=False, use_knn=False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.use_xyz, self.sampling, self.use_knn = use_xyz, sampling, use_knn
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        mlp = list(mlp)
        if self.use_xyz:
            mlp[0] += 3
        last_channel = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.group_all = group_all
        self.self_attention = Self_Attention(last_channel, 2, 'linear')
        self.tc_mode = False      # True: shared MLP on the tensor cores (tcgen05 kind::tf32), "fast" encoder mode
        if group_all or sampling not in ("RANDOM", "FPS") or not use_xyz or len(mlp) != 4:
            raise NotImplementedError("built: sampling 'RANDOM' / 'FPS', use_xyz=True, 3-layer MLP, kNN or ball-query grouping "
                                      "(sample_and_group_edge, pointnet2_utils.py:242-288; backbone_net.py:49-81)")

    def _pack(self):
        # first conv acts on [xyz_j - xyz_c (3), f_c (D), f_j - f_c (D)] (pointnet2_utils.py:279-282):
        #   W [.] = (Wa xyz_j + Wc f_j) + (-Wa xyz_c + (Wb - Wc) f_c)   -> per-point term P1, per-centre term Cc
        w1, b1 = fold_bn(self.mlp_convs[0].weight, self.mlp_convs[0].bias, self.mlp_bns[0])
        D = (w1.shape[1] - 3) // 2
        wa, wb, wc = w1[:, :3], w1[:, 3:3 + D], w1[:, 3 + D:]
        w2, b2 = fold_bn(self.mlp_convs[1].weight, self.mlp_convs[1].bias, self.mlp_bns[1])
        w3, b3 = fold_bn(self.mlp_convs[2].weight, self.mlp_convs[2].bias, self.mlp_bns[2])
        pk = dict(D=D, pa=wa.t().contiguous(), ca=(-wa).t().contiguous(), cbias=b1,
                  w2=w2.t().contiguous(), b2=b2, w3=w3.t().contiguous(), b3=b3)
        if w2.shape[0] in (32, 64, 128):
            pk["w2img"], pk["w3img"] = K.tf32_image(w2), K.tf32_image(w3)
        if D > 0:
            pk["pc"] = wc.t().contiguous()
            pk["cb"] = (wb - wc).t().contiguous()
        return pk

    def forward(self, xyz, points, numpoints):
        """xyz (B, N, 3), points (B, D, N) or None -> new_xyz (B, S, 3), features (B, D', S)."""
        self._inference_only()
        pk = self.packed()
        S = int(numpoints)
        centres = None
        if self.sampling == "FPS":                        # farthest_point_sample (pointnet2_utils.py:116-137), random start
            fps_idx = K.farthest_point_sample(xyz, S)
            new_xyz = torch.gather(xyz, 1, fps_idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
            centres = K.gather_points(points, fps_idx) if points is not None else None
        else:
            new_xyz = xyz[:, :S, :].contiguous()          # sampling == "RANDOM": the first S points
        tc = self.tc_mode and "w2img" in pk          # tensor-core kernel gathers point-major rows
        # (B, S, k) int32; the max over the k edges does not depend on their order: fast mode asks for the set only
        if not self.use_knn:                          # query_ball_point (pointnet2_utils.py:218-240)
            idx = K.query_ball_point(self.radius, self.nsample, xyz, new_xyz)
        else:
            idx = K.knn_point_set(self.nsample, xyz, new_xyz) if tc else K.knn_point(self.nsample, xyz, new_xyz)
        cxyz, cpts = (new_xyz, centres) if self.sampling == "FPS" else (xyz, points)      # centre rows: gathered / the first S
        if pk["D"] > 0:
            p1 = K.cn_linear(xyz, pk["pa"], x2=points, w2=pk["pc"], x1_pm=True, y_pm=tc)
            cc = K.cn_linear(cxyz, pk["ca"], x2=cpts, w2=pk["cb"], bias=pk["cbias"], x1_pm=True, rows=S, y_pm=tc)
        else:
            p1 = K.cn_linear(xyz, pk["pa"], x1_pm=True, y_pm=tc)
            cc = K.cn_linear(cxyz, pk["ca"], bias=pk["cbias"], x1_pm=True, rows=S, y_pm=tc)
        if tc:
            feat = K.sa_edge_mlp_tc(p1, cc, idx, pk["w2img"], pk["b2"], pk["w3img"], pk["b3"])
        else:
            feat = K.sa_edge_mlp(p1, cc, idx, pk["w2"], pk["b2"], pk["w3"], pk["b3"])
        return new_xyz, self.self_attention(feat, new_xyz)


class FP_SA(PackedModule):
    def __init__(self, last_channel, feat1_dim, feat2_dim, d_model, out_dim, nhead, attention='linear'):
        super().__init__()
        self.dim = d_model // nhead
        self.nhead = nhead
        self.pos_mlp2 = nn.Sequential(nn.Linear(3, d_model), nn.ReLU(), nn.Linear(d_model, feat2_dim))
        self.q_proj = nn.Linear(feat1_dim, d_model, bias=False)
        self.k_proj = nn.Linear(feat2_dim, d_model, bias=False)
        self.v_proj = nn.Linear(feat2_dim, d_model, bias=False)
        self.attention = LinearAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(nn.Linear(feat1_dim + d_model, d_model * 2, bias=False), nn.ReLU(True),
                                 nn.Linear(d_model * 2, out_dim, bias=False))
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(out_dim)
        self.tc_mode = False      # True: the block runs as attn_front / linattn_kv_img / attn_back on the tensor cores (fast mode)

    def _pack(self):
        f1 = self.q_proj.weight.shape[1]
        m0 = self.mlp[0].weight.detach()
        pk = dict(
            pos0=kmajor(self.pos_mlp2[0].weight), pos0b=self.pos_mlp2[0].bias.detach().float().contiguous(),
            pos2=kmajor(self.pos_mlp2[2].weight), pos2b=self.pos_mlp2[2].bias.detach().float().contiguous(),
            q=kmajor(self.q_proj.weight), k=kmajor(self.k_proj.weight), v=kmajor(self.v_proj.weight),
            merge=kmajor(self.merge.weight), mlp0a=kmajor(m0[:, :f1]), mlp0b=kmajor(m0[:, f1:]), mlp2=kmajor(self.mlp[2].weight),
            n1w=self.norm1.weight.detach().float().contiguous(), n1b=self.norm1.bias.detach().float().contiguous(),
            n2w=self.norm2.weight.detach().float().contiguous(), n2b=self.norm2.bias.detach().float().contiguous())
        d, f2, out = self.q_proj.weight.shape[0], self.k_proj.weight.shape[1], self.mlp[2].weight.shape[0]
        if fused_attention_supported(d, self.nhead, f1, f2, out):
            f1p = (f1 + 7) // 8 * 8
            pk["front_blob"] = K.weight_blob(K.tf32_image(self.pos_mlp2[2].weight), K.tf32_image(self.v_proj.weight),
                                             K.tf32_image(self.k_proj.weight))
            pk["back_blob"] = K.weight_blob(K.tf32_image_padded(self.q_proj.weight, f1p), K.tf32_image(self.merge.weight),
                                            K.tf32_image_padded(m0[:, :f1], f1p), K.tf32_image(m0[:, f1:]),
                                            K.tf32_image(self.mlp[2].weight))
        return pk

    def forward(self, feat1, xyz1, feat2, xyz2, mask=None, feat1_point_major=False):
        """feat1 (B, C1, N) queries [or point-major (B, N, C1)], feat2 (B, C2, S) keys/values -> (B, out, N)."""
        pk = self.packed()
        S = feat2.shape[2]
        if self.tc_mode and "front_blob" in pk:
            d, out = self.q_proj.weight.shape[0], self.mlp[2].weight.shape[0]
            vk = K.attn_front(xyz2, feat2, pk["pos0"], pk["pos0b"], pk["pos2b"], pk["front_blob"], d, d, d)
            rows_q = feat1.shape[1] if feat1_point_major else feat1.shape[2]
            kvimg, ksum = K.linattn_kv_img(vk[:, d:], vk[:, :d], self.nhead, rows_q)
            return K.attn_back(feat1, None, ksum, kvimg, pk["n1w"], pk["n1b"], pk["n2w"], pk["n2b"], pk["back_blob"],
                               self.nhead, out, residual=False, feat1_pm=feat1_point_major)
        hid = K.cn_linear(xyz2, pk["pos0"], bias=pk["pos0b"], act=K.ACT_RELU, x1_pm=True)
        feat2_pos = K.cn_linear(hid, pk["pos2"], bias=pk["pos2b"], res=feat2)
        q = K.cn_linear(feat1, pk["q"], x1_pm=feat1_point_major)
        k = K.cn_linear(feat2, pk["k"])
        v = K.cn_linear(feat2_pos, pk["v"])
        wkv, ksum = K.linattn_kv(k, v, self.nhead)
        msg = attention_message(q, wkv, ksum, self.nhead, S, pk)
        return attention_ffn(feat1, msg, pk, residual=False, feat_pm=feat1_point_major)


class PointNetFeaturePropagationSA(nn.Module):
    def __init__(self, mlp, mlp_inte):
        super().__init__()
        # dead weights of the reference (never used in its forward, pointnet2_utils.py:441-449, 460-473);
        # kept so that reference checkpoints load with strict=True
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel
        self.interpolation = FP_SA(last_channel=mlp_inte[0], feat1_dim=mlp_inte[1], feat2_dim=mlp_inte[2],
                                   d_model=mlp_inte[3], out_dim=mlp_inte[4], nhead=2, attention='linear')

    def forward(self, xyz1, xyz2, points1, points2, points1_point_major=False):
        return self.interpolation(points1, xyz1, points2, xyz2, feat1_point_major=points1_point_major)
