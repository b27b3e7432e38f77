"""corss_attention -- the pair cross-attention block of the xcorr_eff match head
(mmdet3d/models/attention.py:156-219; the reference's spelling is kept because configs name it), its image-token
variant cross_lin_attn (attention.py:312-372) and local_self_attention (attention.py:221-296)."""
import torch
from torch import nn

from .. import kernels as K
from ._packing import PackedModule, kmajor
from .pointnet2_utils import LinearAttention, attention_ffn, attention_message


class corss_attention(PackedModule):
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

    def _pack(self):
        d = self.q_proj.weight.shape[0]
        m0 = self.mlp[0].weight.detach()
        f = lambda t: t.detach().float().contiguous()
        return dict(
            pos0=kmajor(self.pos_mlp[0].weight), pos0b=f(self.pos_mlp[0].bias),
            pos2=kmajor(self.pos_mlp[2].weight), pos2b=f(self.pos_mlp[2].bias),
            q=kmajor(self.q_proj.weight), k=kmajor(self.k_proj.weight), v=kmajor(self.v_proj.weight),
            merge=kmajor(self.merge.weight), mlp0a=kmajor(m0[:, :d]), mlp0b=kmajor(m0[:, d:]), mlp2=kmajor(self.mlp[2].weight),
            n1w=f(self.norm1.weight), n1b=f(self.norm1.bias), n2w=f(self.norm2.weight), n2b=f(self.norm2.bias))

    # ---- factored entry points used by the all-pairs driver --------------------------------------
    def position_code(self, xyz):
        """pos_mlp(xyz): (B, N, 3) -> (B, C, N).  Depends on the template object only."""
        pk = self.packed()
        hid = K.cn_linear(xyz, pk["pos0"], bias=pk["pos0b"], act=K.ACT_RELU, x1_pm=True)
        return K.cn_linear(hid, pk["pos2"], bias=pk["pos2b"])

    def search_query(self, feat):
        return K.cn_linear(feat, self.packed()["q"])

    def template_summary(self, feat, pos, pos_map=None):
        """-> (Wkv (B, C, C), ksum (B, C)) of a template: k = Wk f, v = Wv (f + pos)."""
        pk = self.packed()
        k = K.cn_linear(feat, pk["k"])
        v = K.cn_linear(feat, pk["v"], x2=pos, w2=pk["v"], x2_map=pos_map)
        return K.linattn_kv(k, v, self.nhead)

    def attend(self, search_feat, q, wkv, ksum, s_len, s_map=None, t_map=None, B=None):
        """search + LN2(mlp(cat[search, LN1(merge(Q.KV.Z))])); maps gather search / template objects per pair."""
        pk = self.packed()
        msg = attention_message(q, wkv, ksum, self.nhead, s_len, pk, q_map=s_map, t_map=t_map, B=B)
        return attention_ffn(search_feat, msg, pk, residual=True, feat_map=s_map)

    # ---- reference API -----------------------------------------------------------------------------
    def forward(self, search_feat, search_xyz, template_feat, template_xyz, mask=None):
        """search_feat (B, C, Ns), search_xyz (B, Ns, 3), template_feat (B, C, Nt), template_xyz (B, Nt, 3) -> (B, C, Ns)."""
        search_feat = _cn_view(search_feat)
        template_feat = _cn_view(template_feat)
        wkv, ksum = self.template_summary(template_feat, self.position_code(template_xyz.contiguous().float()))
        return self.attend(search_feat, self.search_query(search_feat), wkv, ksum, template_feat.shape[2])


class cross_lin_attn(corss_attention):
    """mmdet3d/models/attention.py:312-372 -- the image-token matcher's cross block: corss_attention without the position
    code (`pos_mlp` is constructed, and therefore part of the state_dict, but never used by the reference forward)."""

    def position_code(self, xyz):
        return None

    def template_summary(self, feat, pos=None, pos_map=None):
        """-> (Wkv (B, C, C), ksum (B, C)) of a template token set: k = Wk f, v = Wv f."""
        pk = self.packed()
        return K.linattn_kv(K.cn_linear(feat, pk["k"]), K.cn_linear(feat, pk["v"]), self.nhead)

    def forward(self, search_feat, template_feat, mask=None):
        """search_feat (B, C, Ns), template_feat (B, C, Nt) -> (B, C, Ns)."""
        search_feat = _cn_view(search_feat)
        template_feat = _cn_view(template_feat)
        wkv, ksum = self.template_summary(template_feat)
        return self.attend(search_feat, self.search_query(search_feat), wkv, ksum, template_feat.shape[2])


class local_self_attention(PackedModule):
    """mmdet3d/models/attention.py:221-296: every point attends its `knum` nearest neighbours in FEATURE space
    (the 'xcorr' match type, reid_pts_point-transformer_baseline_orig.py)."""

    def __init__(self, d_model, nhead, attention='linear', knum=32, pos_size=16):
        super().__init__()
        self.d_model = d_model
        self.dim = d_model // nhead
        self.nhead = nhead
        self.knum = knum
        self.pos_mlp_knn = nn.Sequential(nn.Linear(3, pos_size), nn.ReLU(True), nn.Linear(pos_size, pos_size))
        self.q_proj_knn = nn.Linear(d_model, d_model, bias=False)
        self.k_proj_knn = nn.Linear(d_model, d_model, bias=False)
        self.v_proj_knn = nn.Linear(d_model, d_model, bias=False)
        self.attention_knn = LinearAttention()
        self.merge_knn = nn.Linear(d_model, d_model, bias=False)
        self.mlp_knn = nn.Sequential(nn.Linear(d_model * 2, d_model * 2, bias=False), nn.ReLU(True),
                                     nn.Linear(d_model * 2, d_model, bias=False))
        self.norm1_knn = nn.LayerNorm(d_model)
        self.norm2_knn = nn.LayerNorm(d_model)
        if pos_size != d_model:
            raise ValueError("local_self_attention adds pos_mlp_knn(xyz) to the features: pos_size must equal d_model "
                             "(attention.py:277-278)")

    def _pack(self):
        d = self.d_model
        m0 = self.mlp_knn[0].weight.detach()
        f = lambda t: t.detach().float().contiguous()
        return dict(
            pos0=kmajor(self.pos_mlp_knn[0].weight), pos0b=f(self.pos_mlp_knn[0].bias),
            pos2=kmajor(self.pos_mlp_knn[2].weight), pos2b=f(self.pos_mlp_knn[2].bias),
            qkv=kmajor(torch.cat([self.q_proj_knn.weight, self.k_proj_knn.weight, self.v_proj_knn.weight], 0)),
            merge=kmajor(self.merge_knn.weight), mlp0a=kmajor(m0[:, :d]), mlp0b=kmajor(m0[:, d:]), mlp2=kmajor(self.mlp_knn[2].weight),
            n1w=f(self.norm1_knn.weight), n1b=f(self.norm1_knn.bias), n2w=f(self.norm2_knn.weight), n2b=f(self.norm2_knn.bias))

    def forward(self, search_feat, search_xyz, mask=None):
        """search_feat (B, C, N), search_xyz (B, N, 3) -> (B, C, N)."""
        pk = self.packed()
        feat = _cn_view(search_feat)
        if not feat.is_contiguous():
            feat = feat.contiguous()
        xyz = search_xyz.contiguous().float()
        kidx = K.knn_feature(feat, self.knum)                                   # (B, N, knum), includes the point itself
        hid = K.cn_linear(xyz, pk["pos0"], bias=pk["pos0b"], act=K.ACT_RELU, x1_pm=True)
        feat_pos = K.cn_linear(hid, pk["pos2"], bias=pk["pos2b"], res=feat)
        qkv = K.cn_linear(feat_pos, pk["qkv"], y_pm=True)                       # (B, N, 3C) point-major rows
        att = K.local_linattn(qkv, kidx, self.nhead)                            # (B, N, C) point-major
        msg = K.cn_groupnorm(K.cn_linear(att, pk["merge"], x1_pm=True), pk["n1w"], pk["n1b"], 1)
        return attention_ffn(feat, msg, pk, residual=True)


def _cn_view(t):
    """accepts any (B, C, N) float tensor; copies only if the point axis is not unit-stride."""
    t = t.float()
    return t if t.stride(2) == 1 or t.shape[2] == 1 else t.contiguous()
