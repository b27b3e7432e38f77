"""TEST-ONLY torch/CPU emulation of the C-ABI kernels' *specification* (include/pcreid.h), used to check
the host-side logic (weight packing, BatchNorm folding, conv factorisation, object maps, layouts) of
the model modules against the oracle on machines without a GPU.  Never imported by the product."""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_LEAKY02, ACT_ELU1 = 0, 1, 2, 3


def _act(x, act):
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_LEAKY02:
        return F.leaky_relu(x, 0.2)
    if act == ACT_ELU1:
        return F.elu(x) + 1
    return x


def _gather(x, m):
    return x if m is None else x[m.long()]


def cn_linear(x1, w1, x2=None, w2=None, bias=None, act=0, res=None, res_after_act=False, rows=None, x1_map=None,
              x2_map=None, w1_map=None, r_map=None, x1_pm=False, x2_pm=False, out=None, B=None, y_pm=False):
    a = _gather(x1, x1_map)
    a = a.transpose(1, 2) if x1_pm else a                       # (B, K, N)
    rows = a.shape[2] if rows is None else rows
    a = a[:, :, :rows]
    w = w1 if w1.dim() == 2 else _gather(w1, w1_map)
    if B is not None and a.shape[0] == 1 and B > 1:
        a = a.expand(B, -1, -1)
    y = torch.einsum("bkn,kc->bcn", a, w) if w.dim() == 2 else torch.einsum("bkn,bkc->bcn", a, w)
    if x2 is not None:
        b_ = _gather(x2, x2_map)
        b_ = (b_.transpose(1, 2) if x2_pm else b_)[:, :, :rows]
        y = y + (torch.einsum("bkn,kc->bcn", b_, w2) if w2.dim() == 2 else torch.einsum("bkn,bkc->bcn", b_, w2))
    if bias is not None:
        y = y + bias.view(1, -1, 1)
    r = None if res is None else _gather(res, r_map)[:, :, :rows]
    if r is not None and not res_after_act:
        y = y + r
    y = _act(y, act)
    if r is not None and res_after_act:
        y = y + r
    if y_pm:
        y = y.transpose(1, 2).contiguous()
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def cn_groupnorm(x, gamma, beta, groups=1, res=None, r_map=None, act=0, out=None):
    B, C, N = x.shape
    xr = x.permute(0, 2, 1).reshape(B * N, C)
    y = F.group_norm(xr, groups, gamma, beta, 1e-5).reshape(B, N, C).permute(0, 2, 1)
    if res is not None:
        y = y + _gather(res, r_map)
    return _act(y, act).contiguous()


def linattn_kv(k, v, nhead):
    B, d, S = k.shape
    D = d // nhead
    Kf = (F.elu(k) + 1).view(B, nhead, D, S)
    V = (v / S).view(B, nhead, D, S)
    kv = torch.einsum("bhis,bhjs->bhij", Kf, V)
    w = torch.zeros(B, d, d)
    for h in range(nhead):
        w[:, h * D:(h + 1) * D, h * D:(h + 1) * D] = kv[:, h]
    return w, Kf.sum(-1).reshape(B, d)


def linattn_scale(q, ksum, nhead, s_len, q_map=None, ksum_map=None, B=None):
    q = _gather(q, q_map)
    ks = _gather(ksum, ksum_map)
    Bq, d, N = q.shape
    D = d // nhead
    Q = (F.elu(q) + 1).view(Bq, nhead, D, N)
    z = 1.0 / (torch.einsum("bhdn,bhd->bhn", Q, ks.view(-1, nhead, D)) + 1e-6) * s_len
    return (Q * z.unsqueeze(2)).reshape(Bq, d, N).contiguous()


def cn_pool(x1, x2=None, mode=0, out=None, transposed=False):
    x = x1 if x2 is None else torch.cat([x1, x2], 2)
    o = x.max(2)[0] if mode == 1 else torch.cat([x.max(2)[0], x.mean(2)], 1)
    return o.t().contiguous().unsqueeze(0) if transposed else o.contiguous()


def cn_chanmax(x, out=None, transposed=False):
    o = x.max(1)[0]
    return o.t().contiguous().unsqueeze(0) if transposed else o.contiguous()


def knn_point(k, xyz, new_xyz):
    d = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    d += torch.sum(new_xyz ** 2, -1).unsqueeze(-1)
    d += torch.sum(xyz ** 2, -1).unsqueeze(1)
    return torch.sort(d, dim=-1, stable=True)[1][..., :k].to(torch.int32).contiguous()


def farthest_point_sample(xyz, npoint, start=None):
    """specification of pcreid_fps_torch: running minimum of (dx*dx + dy*dy) + dz*dz, first index among tied maxima."""
    B, N, _ = xyz.shape
    far = torch.randint(0, N, (B,), dtype=torch.long) if start is None else start.long().clone()
    dist = torch.full((B, N), 1e10)
    out = torch.zeros(B, npoint, dtype=torch.int32)
    for i in range(npoint):
        out[:, i] = far.to(torch.int32)
        c = xyz[torch.arange(B), far].view(B, 1, 3)
        d = xyz - c
        dist = torch.minimum(dist, (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
        far = (dist == dist.max(-1, keepdim=True)[0]).float().argmax(-1)
    return out


def gather_points(features, idx):
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1)).contiguous()


def query_ball_point(radius, nsample, xyz, new_xyz):
    """specification of pcreid_query_ball_point: first nsample indices with d <= r^2 (fp32), padded with the first; N if none."""
    B, N, _ = xyz.shape
    d = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    d += torch.sum(new_xyz ** 2, -1).unsqueeze(-1)
    d += torch.sum(xyz ** 2, -1).unsqueeze(1)
    keep = ~(d > float(torch.tensor(radius ** 2, dtype=torch.float32)))
    ar = torch.arange(N).view(1, 1, N).expand_as(d)
    cand = torch.where(keep, ar, torch.full_like(ar, N)).sort(-1)[0][..., :nsample]
    first = cand[..., :1].expand_as(cand)
    return torch.where(cand == N, first, cand).to(torch.int32).contiguous()


def knn_feature(x, k):
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return torch.sort(pd, dim=-1, descending=True, stable=True)[1][..., :k].to(torch.int32).contiguous()


def local_linattn(qkv_pm, idx, nhead):
    """specification of pcreid_local_linattn: one query per point over its gathered neighbours."""
    B, N, C3 = qkv_pm.shape
    C, k = C3 // 3, idx.shape[2]
    q = F.elu(qkv_pm[..., :C]) + 1
    bidx = torch.arange(B).view(B, 1, 1).expand(B, N, k)
    nb = qkv_pm[bidx, idx.long()]                                            # (B, N, k, 3C)
    kf, v = F.elu(nb[..., C:2 * C]) + 1, nb[..., 2 * C:]
    dh = C // nhead
    w = (q.view(B, N, 1, nhead, dh) * kf.view(B, N, k, nhead, dh)).sum(-1)   # (B, N, k, H)
    num = (w.unsqueeze(-1) * v.view(B, N, k, nhead, dh)).sum(2)              # (B, N, H, dh)
    return (num / (w.sum(2).unsqueeze(-1) + 1e-6)).reshape(B, N, C).contiguous()


def _gather_pts(p, idx):
    B, C, N = p.shape
    S, k = idx.shape[1], idx.shape[2]
    return torch.gather(p.unsqueeze(2).expand(B, C, S, N), 3, idx.long().unsqueeze(1).expand(B, C, S, k))


def sa_edge_mlp(p1, cc, idx, w2, b2, w3, b3):
    h = F.relu(_gather_pts(p1, idx) + cc.unsqueeze(-1))                  # (B, C, S, k)
    h = F.relu(torch.einsum("bcsk,cd->bdsk", h, w2) + b2.view(1, -1, 1, 1))
    h = F.relu(torch.einsum("bcsk,cd->bdsk", h, w3) + b3.view(1, -1, 1, 1))
    return h.max(-1)[0].contiguous()


def tf32_image(w):
    n, kd = w.shape
    return w.detach().float().reshape(n, kd // 4, 4).permute(1, 0, 2).contiguous()


def sa_edge_mlp_tc(p1, cc, idx, w2img, b2, w3img, b3):
    C = p1.shape[2]                                                            # p1 (B, N, C), cc (B, S, C) point-major
    unimg = lambda im: im.permute(1, 0, 2).reshape(C, C).t().contiguous()      # -> k-major (C_in, C_out)
    return sa_edge_mlp(p1.transpose(1, 2).contiguous(), cc.transpose(1, 2).contiguous(), idx, unimg(w2img), b2, unimg(w3img), b3)


def edge_gather_max(p, q, idx, act, out=None):
    y = _act(_gather_pts(p, idx).max(-1)[0] + q, act)
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def pair_concat_head(a, bv, et, ed, w2, g1, be1, g2, be2, w, b0, groups, mask=None):
    T, D = a.shape[0], bv.shape[0]
    x = (a[:, None, :] + bv[None, :, :]).reshape(T * D, -1)
    h = F.relu(F.group_norm(x, groups, g1, be1, 1e-5))
    y = F.group_norm(h @ w2, groups, g2, be2, 1e-5)
    r = torch.cat([et[:, None, :].expand(T, D, -1), ed[None, :, :].expand(T, D, -1)], -1).reshape(T * D, -1)
    o = (F.relu(y + r) @ w + b0).reshape(T, D)
    if mask is not None:
        o = o * (mask != 0)
    return o


def install(monkeypatch=None):
    """Replaces pcreid_b200.kernels' entry points by the emulations above (optionally via pytest's monkeypatch)."""
    import pcreid_b200.kernels as K
    names = ["cn_linear", "cn_groupnorm", "linattn_kv", "linattn_scale", "cn_pool", "cn_chanmax", "knn_point",
             "query_ball_point", "farthest_point_sample", "gather_points", "knn_feature", "sa_edge_mlp", "sa_edge_mlp_tc", "tf32_image", "edge_gather_max", "pair_concat_head", "local_linattn"]
    for n in names:
        if monkeypatch is not None:
            monkeypatch.setattr(K, n, globals()[n])
        else:
            setattr(K, n, globals()[n])
